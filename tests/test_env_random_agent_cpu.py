"""Random-agent fuzz of the drop-in boundary, CPU: two unmodified grid2op environments side by side — ``B200Backend`` (its HOST logic:
apply_action / topology bookkeeping / result slicing / copy / reset, engine replaced by the oracle adapter = test infrastructure) and
the oracle's restatement of ``PandaPowerBackend`` — fed the same ``action_space.sample()`` stream (set_bus / change_bus / line status
/ redispatch / curtailment / storage, whatever the environment's action class allows) with the same seeds.  After every step the
observations must agree (integers exactly), game overs must coincide; after a game over both reset and go on.
Reference: the sequence ``BaseEnv.step`` -> ``_backend_action += action`` -> ``apply_action_public`` -> ``next_grid_state``
(grid2op/Environment/baseEnv.py:3778-3872, grid2op/Backend/backend.py:450-496, :1433-1521)."""
import os
import warnings

import numpy as np
import pytest

from conftest import env_grid


def _close(a, b, tol, what):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    assert np.array_equal(np.isfinite(a), np.isfinite(b)), what
    m = np.isfinite(a)
    if m.any():
        err = float(np.max(np.abs(a[m] - b[m])))
        assert err <= tol, (what, err)


def _compare(o1, o2, sn_mva, step):
    tol_mw = 1e-4 * sn_mva + 4e-6 * 2000.0
    for k in ("p_or", "q_or", "p_ex", "q_ex", "gen_p", "gen_q", "load_p", "load_q", "storage_power", "actual_dispatch", "target_dispatch"):
        _close(getattr(o1, k), getattr(o2, k), tol_mw, (k, step))
    vmax = max(1.0, float(np.nanmax(np.abs(o2.v_or))))
    for k in ("v_or", "v_ex", "gen_v", "load_v"):
        _close(getattr(o1, k), getattr(o2, k), 1e-4 * vmax, (k, step))
    _close(o1.rho, o2.rho, 2e-4, ("rho", step))
    _close(o1.storage_charge, o2.storage_charge, 1e-3, ("storage_charge", step))
    for k in ("theta_or", "theta_ex", "gen_theta", "load_theta"):
        _close(getattr(o1, k), getattr(o2, k), 2e-3, (k, step))
    _close(o1.storage_theta, o2.storage_theta, 2e-3, ("storage_theta", step))      # (the reference's "theta" of a storage unit is a voltage)
    for k in ("topo_vect", "line_status", "timestep_overflow", "time_before_cooldown_line", "time_before_cooldown_sub",
              "time_next_maintenance", "duration_next_maintenance"):
        assert np.array_equal(getattr(o1, k), getattr(o2, k)), (k, step)


# (environment, steps, sn_mva, DC environment, busbars per substation, obs.simulate every third step)
CASES = [("l2rpn_case14_sandbox", 120, 100.0, False, 2, True), ("educ_case14_storage", 120, 100.0, False, 2, False),
         ("l2rpn_neurips_2020_track1", 60, 1.0, False, 2, False), ("l2rpn_wcci_2022_dev", 30, 1.0, False, 2, False),
         # DC environments (Parameters.ENV_DC): q = 0, load_v conventions, and the reference switching its storage units off after a
         # DC solve (pPB:1203-1207)
         ("educ_case14_storage", 60, 100.0, True, 2, False), ("l2rpn_wcci_2022_dev", 16, 1.0, True, 2, False),
         ("rte_case5_example", 60, 1.0, True, 2, True),
         # three busbars per substation (grid2op.make(..., n_busbar=3))
         ("l2rpn_case14_sandbox", 60, 100.0, False, 3, False), ("l2rpn_neurips_2020_track1", 40, 1.0, False, 3, False)]


@pytest.mark.parametrize("name,n_steps,sn_mva,dc,n_busbar,with_simulate", CASES)
def test_random_agent_side_by_side(name, n_steps, sn_mva, dc, n_busbar, with_simulate):
    if env_grid(name) is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk           # (locates / bootstraps the grid2op install first)
    from oracle_engine import OracleEngine

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    run_side_by_side(HostLogicBackend, name, n_steps, sn_mva, dc, n_busbar, with_simulate)


def run_side_by_side(backend_class, name, n_steps, sn_mva, dc, n_busbar, with_simulate, tag="fuzz"):
    """(also the body of the GPU variant, tests/test_zz_random_agents_gpu.py, with the CUDA backend as ``backend_class``)"""
    import grid2op_b200.backend as bk           # noqa: F401  (locates / bootstraps the grid2op install first)
    from grid2op.Parameters import Parameters
    param = Parameters()
    param.ENV_DC = bool(dc)
    kw = {"param": param}
    if n_busbar != 2:
        kw["n_busbar"] = n_busbar
    import grid2op
    from oracle.ppbackend_ref import PandaPowerBackendRef
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e1 = grid2op.make(name, test=True, backend=backend_class(), _add_to_name=f"{tag}_b200_{n_busbar}", **kw)
        e2 = grid2op.make(name, test=True, backend=PandaPowerBackendRef(), _add_to_name=f"{tag}_ref_{n_busbar}", **kw)
    try:
        for e in (e1, e2):
            e.seed(3); e.set_id(0)
        o1, o2 = e1.reset(), e2.reset()
        _compare(o1, o2, sn_mva, "reset")
        e1.action_space.seed(11); e2.action_space.seed(11)
        n_over = n_acted = 0
        for i in range(n_steps):
            # every other step a do-nothing, so that the grid lives long enough for cooldowns / storage / dispatch to matter
            a1 = e1.action_space.sample() if i % 2 == 0 else e1.action_space()
            a2 = e2.action_space.sample() if i % 2 == 0 else e2.action_space()
            assert np.array_equal(a1.to_vect(), a2.to_vect(), equal_nan=True)
            if with_simulate and i % 3 == 0:             # what-if first: Backend.copy() + the forecast injections
                s1, _, ds1, _ = o1.simulate(a1)
                s2, _, ds2, _ = o2.simulate(a2)
                assert ds1 == ds2, ("simulate done", i)
                if not ds1:
                    _compare(s1, s2, sn_mva, ("simulate", i))
            o1, r1, d1, i1 = e1.step(a1)
            o2, r2, d2, i2 = e2.step(a2)
            assert d1 == d2, (i, i1["exception"], i2["exception"])
            assert i1["is_illegal"] == i2["is_illegal"] and i1["is_ambiguous"] == i2["is_ambiguous"], i
            assert np.array_equal(i1["disc_lines"], i2["disc_lines"]), i
            if d1:
                n_over += 1
                o1, o2 = e1.reset(), e2.reset()
                _compare(o1, o2, sn_mva, ("reset after game over", i))
                continue
            n_acted += 1
            _compare(o1, o2, sn_mva, i)
            assert abs(r1 - r2) <= 1e-3 * max(1.0, abs(r2)), (i, r1, r2)
        assert n_acted >= n_steps // 5, (n_acted, n_over)
    finally:
        e1.close(); e2.close()


@pytest.mark.parametrize("name,sn_mva,dc", [("l2rpn_case14_sandbox", 100.0, False), ("l2rpn_case14_sandbox", 100.0, True),
                                            ("educ_case14_storage", 100.0, False), ("educ_case14_storage", 100.0, True),
                                            ("rte_case5_example", 1.0, False)])
def test_random_element_actions_with_detachment(name, sn_mva, dc):
    """Same side-by-side run with the actions ``action_space.sample()`` never draws: loads / generators / storage units moved between
    busbars or DETACHED (``allow_detachment=True``, set_bus -1: grid2op/Backend/backend.py:262-329), shunts switched, moved and
    re-set (``shunt`` key), in AC and DC environments; the shunt part of the observation is compared too."""
    if env_grid(name) is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    from grid2op.Parameters import Parameters
    from oracle_engine import OracleEngine
    import grid2op
    from oracle.ppbackend_ref import PandaPowerBackendRef

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    p = Parameters()
    p.ENV_DC = bool(dc)
    p.MAX_SUB_CHANGED = 99
    p.MAX_LINE_STATUS_CHANGED = 99
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from grid2op.Action import CompleteAction        # (every key allowed, whatever the environment's own action class is)
        e1 = grid2op.make(name, test=True, param=p, allow_detachment=True, action_class=CompleteAction, backend=HostLogicBackend(),
                          _add_to_name="det_b200")
        e2 = grid2op.make(name, test=True, param=p, allow_detachment=True, action_class=CompleteAction, backend=PandaPowerBackendRef(),
                          _add_to_name="det_ref")
    rng = np.random.default_rng(4)
    try:
        for e in (e1, e2):
            e.seed(1); e.set_id(0)
        o1, o2 = e1.reset(), e2.reset()
        _compare(o1, o2, sn_mva, "reset")
        n_ok = 0
        for i in range(50):
            r = rng.random()
            spec = {}
            if r < 0.3:
                spec = {"set_bus": {"loads_id": [(int(rng.integers(0, e1.n_load)), int(rng.choice([-1, 1, 2])))]}}
            elif r < 0.5:
                spec = {"set_bus": {"generators_id": [(int(rng.integers(0, e1.n_gen)), int(rng.choice([-1, 1, 2])))]}}
            elif r < 0.6 and e1.n_storage:
                spec = {"set_bus": {"storages_id": [(int(rng.integers(0, e1.n_storage)), int(rng.choice([-1, 1, 2])))]}}
            elif r < 0.7 and e1.n_shunt:
                spec = {"shunt": {"set_bus": [(int(rng.integers(0, e1.n_shunt)), int(rng.choice([-1, 1, 2])))]}}
            elif r < 0.8 and e1.n_shunt:
                spec = {"shunt": {"shunt_q": [(int(rng.integers(0, e1.n_shunt)), float(rng.uniform(-30, 30)))]}}
            o1, r1, d1, i1 = e1.step(e1.action_space(spec))
            o2, r2, d2, i2 = e2.step(e2.action_space(spec))
            assert d1 == d2, (i, spec, i1["exception"], i2["exception"])
            if d1:
                o1, o2 = e1.reset(), e2.reset()
                _compare(o1, o2, sn_mva, ("reset after game over", i))
                continue
            n_ok += 1
            _compare(o1, o2, sn_mva, (i, spec))
            for k in ("_shunt_p", "_shunt_q", "_shunt_v", "_shunt_bus"):
                assert np.allclose(getattr(o1, k), getattr(o2, k), atol=1e-3, equal_nan=True), (k, i, spec)
        assert n_ok >= 10
    finally:
        e1.close(); e2.close()


@pytest.mark.parametrize("name", ["educ_case14_redisp", "l2rpn_case14_sandbox_diff_grid", "l2rpn_icaps_2021", "l2rpn_idf_2023",
                                  "l2rpn_neurips_2020_track2", "l2rpn_wcci_2020", "rte_case118_example", "rte_case14_opponent",
                                  "rte_case14_realistic", "rte_case14_redisp", "rte_case14_test"])
def test_every_other_bundled_environment(name):
    """A short sampled-action run on every remaining environment the reference bundles (grid2op/data/*: other grid files, trafo
    parameters, opponents, multi-mix, 118-substation grids): same observations as the oracle backend."""
    if env_grid(name) is None and not os.path.isdir(os.path.join(os.path.dirname(os.path.dirname(env_grid("l2rpn_case14_sandbox") or "/x/y")), name)):
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    from oracle_engine import OracleEngine
    import grid2op
    from oracle.ppbackend_ref import PandaPowerBackendRef

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e1 = grid2op.make(name, test=True, backend=HostLogicBackend(), _add_to_name="all_b200")
        e2 = grid2op.make(name, test=True, backend=PandaPowerBackendRef(), _add_to_name="all_ref")
    try:
        for e in (e1, e2):
            e.seed(0); e.set_id(0)
        o1, o2 = e1.reset(), e2.reset()
        _compare(o1, o2, 100.0, "reset")
        e1.action_space.seed(7); e2.action_space.seed(7)
        for i in range(8):
            a1 = e1.action_space.sample() if i % 2 == 0 else e1.action_space()
            a2 = e2.action_space.sample() if i % 2 == 0 else e2.action_space()
            o1, r1, d1, i1 = e1.step(a1)
            o2, r2, d2, i2 = e2.step(a2)
            assert d1 == d2, (i, i1["exception"], i2["exception"])
            if d1:
                o1, o2 = e1.reset(), e2.reset()
            _compare(o1, o2, 100.0, i)
    finally:
        e1.close(); e2.close()


def _getters(b):
    out = {"topo": b.get_topo_vect(), "status": b.get_line_status()}
    for nm, fn in (("gen", b.generators_info), ("load", b.loads_info), ("or", b.lines_or_info), ("ex", b.lines_ex_info),
                   ("storage", b.storages_info), ("shunt", b.shunt_info)):
        for k, a in enumerate(fn()):
            out[f"{nm}{k}"] = np.asarray(a, dtype=np.float64)
    for k, a in enumerate(b.get_theta()):
        out[f"theta{k}"] = np.asarray(a, dtype=np.float64)
    return out


@pytest.mark.parametrize("name", ["l2rpn_case14_sandbox", "educ_case14_storage", "l2rpn_neurips_2020_track1"])
def test_backend_level_sequences_without_an_environment(name):
    """The Backend API driven directly (the way aaa_test_backend_interface.py drives it): random sequences of ``apply_action`` with
    sampled TOPOLOGY actions on a backend action, ``copy()``, and ``runpf`` in AC or DC in any order — every getter of ``B200Backend``
    (host logic) equal to the oracle backend's after every converged solve.  (Injection-type actions are left out: a fresh
    ``_BackendAction`` holds uninitialised set points, ``ValueStore`` is ``np.empty``.)"""
    if env_grid(name) is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    from oracle_engine import OracleEngine
    import grid2op
    from grid2op.Action import CompleteAction
    from oracle.ppbackend_ref import PandaPowerBackendRef

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e1 = grid2op.make(name, test=True, action_class=CompleteAction, backend=HostLogicBackend(), _add_to_name="bkseq_b200")
        e2 = grid2op.make(name, test=True, action_class=CompleteAction, backend=PandaPowerBackendRef(), _add_to_name="bkseq_ref")
    try:
        n_conv = n_dc = n_fail = 0
        for seed in range(2):
            e1.action_space.seed(seed); e2.action_space.seed(seed)
            b1, b2 = e1.backend.copy(), e2.backend.copy()
            ba1, ba2 = type(b1).my_bk_act_class(), type(b2).my_bk_act_class()
            rng = np.random.default_rng(seed)
            for i in range(40):
                r = rng.random()
                if r < 0.5:
                    a1, a2 = e1.action_space.sample(), e2.action_space.sample()
                    while a1._modif_inj or a1._modif_redispatch or a1._modif_storage or a1._modif_curtailment or a1._modif_alarm or a1._modif_alert:
                        a1, a2 = e1.action_space.sample(), e2.action_space.sample()
                    ba1 += a1; ba2 += a2
                    b1.apply_action(ba1); b2.apply_action(ba2)
                    ba1.reset(); ba2.reset()
                elif r < 0.6:
                    b1, b2 = b1.copy(), b2.copy()
                dc = bool(rng.random() < 0.3)
                c1, x1 = b1.runpf(is_dc=dc)
                c2, x2 = b2.runpf(is_dc=dc)
                assert c1 == c2, (seed, i, dc, x1, x2)
                if not c1:                      # like an environment after a game over: back to the loaded state
                    n_fail += 1
                    b1, b2 = e1.backend.copy(), e2.backend.copy()
                    ba1, ba2 = type(b1).my_bk_act_class(), type(b2).my_bk_act_class()
                    continue
                n_conv += 1; n_dc += int(dc)
                g1, g2 = _getters(b1), _getters(b2)
                for k in g1:
                    _close(g1[k], g2[k], 2e-3 * max(1.0, 0.1 * float(np.nanmax(np.abs(g2[k]), initial=0.0))), (k, seed, i, dc))
        assert n_conv >= 20 and n_dc >= 3 and n_fail >= 1, (n_conv, n_dc, n_fail)
    finally:
        e1.close(); e2.close()
