"""Two unmodified grid2op environments side by side on l2rpn_case14_sandbox: one with B200Backend (CUDA), one
with the oracle's restatement of PandaPowerBackend (CPU).  Same seeds, same actions (topology changes, line
disconnection / reconnection, DoNothing), plus obs.simulate() and env.copy(): observations must agree within
the north_star tolerance (1e-4 p.u. on flows and voltages; integers exactly)."""
import warnings

import numpy as np
import pytest

from conftest import env_grid

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    m = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(np.isfinite(a), np.isfinite(b)), what
    if m.any():
        assert np.max(np.abs(a[m] - b[m])) <= tol, (what, float(np.max(np.abs(a[m] - b[m]))))


WITH_THETA = False      # the CPU counterpart (and the never-run GPU variant in test_zz_random_agents_gpu.py) switch the angle comparison on


def _compare_obs(o1, o2, sn_mva, vn_max):
    tol_mw = 1e-4 * sn_mva + 4e-6 * 300.0        # 1e-4 p.u. + float32 resolution of the observation
    for k in ("p_or", "q_or", "p_ex", "q_ex", "gen_p", "gen_q", "load_p", "load_q"):
        _close(getattr(o1, k), getattr(o2, k), tol_mw, k)
    for k in ("v_or", "v_ex", "gen_v", "load_v"):
        _close(getattr(o1, k), getattr(o2, k), 1e-4 * vn_max, k)
    _close(o1.rho, o2.rho, 1e-4, "rho")
    # angles, incl. the reference's quirk for OPEN lines (the angle of the bus the end was last attached to, pPB:1163-1187)
    if WITH_THETA:
        for k in ("theta_or", "theta_ex", "gen_theta", "load_theta"):
            _close(getattr(o1, k), getattr(o2, k), 1e-2, k)
    for k in ("topo_vect", "line_status", "timestep_overflow"):
        assert np.array_equal(getattr(o1, k), getattr(o2, k)), k


def test_case14_sandbox_env_side_by_side(cuda_required):
    if env_grid("l2rpn_case14_sandbox") is None:
        pytest.skip("reference data not available")
    from grid2op_b200.backend import B200Backend          # locates / bootstraps the grid2op install first
    import grid2op
    from oracle.ppbackend_ref import PandaPowerBackendRef
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e1 = grid2op.make("l2rpn_case14_sandbox", test=True, backend=B200Backend(), _add_to_name="sbs_b200")
        e2 = grid2op.make("l2rpn_case14_sandbox", test=True, backend=PandaPowerBackendRef(), _add_to_name="sbs_ref")
    for e in (e1, e2):
        e.seed(0); e.set_id(0)
    o1, o2 = e1.reset(), e2.reset()
    _compare_obs(o1, o2, 100.0, 138.0)
    acts = [
        {},
        {"set_bus": {"substations_id": [(1, [1, 2, 1, 2, 1, 2])]}},
        {},
        {"set_line_status": [(13, -1)]},
        {},
        {"set_bus": {"substations_id": [(4, [1, 1, 2, 2, 1])]}},
        {"set_line_status": [(13, +1)]},
        {},
        {"set_bus": {"substations_id": [(1, [1, 1, 1, 1, 1, 1])]}},
        {},
    ]
    for i, spec in enumerate(acts):
        a1, a2 = e1.action_space(spec), e2.action_space(spec)
        # what-if on the current observation first (uses Backend.copy() and the forecasted injections)
        s1, _, d1s, i1s = o1.simulate(a1)
        s2, _, d2s, i2s = o2.simulate(a2)
        assert d1s == d2s, ("simulate done", i)
        if not d1s:
            _compare_obs(s1, s2, 100.0, 138.0)
        o1, r1, d1, info1 = e1.step(a1)
        o2, r2, d2, info2 = e2.step(a2)
        assert d1 == d2, ("done", i, info1["exception"], info2["exception"])
        assert not d1, (i, info1["exception"])
        _compare_obs(o1, o2, 100.0, 138.0)
        assert abs(r1 - r2) <= 1e-2 * max(1.0, abs(r2))
    # a copied environment continues identically
    c1, c2 = e1.copy(), e2.copy()
    for _ in range(3):
        oc1, *_ = c1.step(c1.action_space())
        oc2, *_ = c2.step(c2.action_space())
        _compare_obs(oc1, oc2, 100.0, 138.0)
    for e in (e1, e2, c1, c2):
        e.close()
