"""Kirchhoff pin for the grid whose file holds no stored pandapower result (l2rpn_case14_sandbox — the headline grid):
SURVEY.md 8(c) names KCL residuals as the fallback pin where no reference numbers exist.

1. ``Backend.check_kirchhoff`` (reference grid2op/Backend/backend.py:1576-1666; tolerance 1e-2 MW / MVAr of
   grid2op/tests/BaseBackendTest.py:843-859) after EVERY step of DoNothing rollouts over all bundled chronics rows of the
   environment (3 scenarios x 575 steps) through the unmodified grid2op Environment + B200Backend (CUDA);
2. the same balance at batch scale: 4096 instances stepped in series mode, per-bus active / reactive sums rebuilt in numpy
   from the packed result records (validated against check_kirchhoff's own numbers in 1).
The actual residuals are written to gpurun_out/kcl_case14.json (-> profiles/)."""
import json
import os
import warnings

import numpy as np
import pytest

from conftest import env_grid

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kcl_residual(gm, out, rows):
    """per-substation sums of active / reactive power (MW / MVAr, everything on busbar 1): lines leaving + loads - units + shunts"""
    from grid2op_b200.engine import OutputView
    v = OutputView(gm, out)
    B = out.shape[0]
    p = np.zeros((B, gm.n_sub)); q = np.zeros((B, gm.n_sub))
    nl = gm.n_load
    for arr_p, arr_q, sub, sign in ((v.p_or, v.q_or, gm.line_or_sub, 1.0), (v.p_ex, v.q_ex, gm.line_ex_sub, 1.0),
                                    (rows[:, :nl], rows[:, nl:2 * nl], gm.load_sub, 1.0),
                                    (v.unit_p, v.unit_q, gm.unit_sub, -1.0), (v.shunt_p, v.shunt_q, gm.shunt_sub, 1.0)):
        for k, s in enumerate(np.asarray(sub)):
            p[:, s] += sign * arr_p[:, k].astype(np.float64)
            q[:, s] += sign * arr_q[:, k].astype(np.float64)
    return p, q


def _dump(name, payload):
    d = os.path.join(REPO, "gpurun_out")
    if os.path.isdir(d):
        path = os.path.join(d, "kcl_case14.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[name] = payload
        json.dump(cur, open(path, "w"), indent=1)


def test_check_kirchhoff_over_all_chronics_rows_through_the_env(cuda_required):
    if env_grid("l2rpn_case14_sandbox") is None:
        pytest.skip("reference data not available")
    from grid2op_b200.backend import B200Backend
    import grid2op
    from grid2op.Parameters import Parameters
    p = Parameters()
    p.NO_OVERFLOW_DISCONNECTION = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = grid2op.make("l2rpn_case14_sandbox", test=True, backend=B200Backend(), param=p, _add_to_name="kcl")
    worst = np.zeros(4)
    n = 0
    for sc in range(3):
        env.set_id(sc)
        env.reset()
        done = False
        while not done:
            _, _, done, info = env.step(env.action_space())
            if done:
                break
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                p_subs, q_subs, p_bus, q_bus, dv = env.backend.check_kirchhoff()
            worst = np.maximum(worst, [np.abs(p_subs).max(), np.abs(q_subs).max(), np.abs(p_bus).max(), np.abs(q_bus).max()])
            n += 1
    env.close()
    assert n >= 1700, n
    assert worst.max() <= 1e-2, worst                    # the reference's tolerance (MW / MVAr)
    _dump("env_rollouts", {"steps": n, "max_abs_p_sub_MW": worst[0], "max_abs_q_sub_MVAr": worst[1], "max_abs_p_bus_MW": worst[2],
                           "max_abs_q_bus_MVAr": worst[3], "max_pu_sn100": float(worst.max() / 100.0), "tolerance_MW": 1e-2})
    assert worst.max() / 100.0 <= 1e-6                   # and in p.u. of sn_mva = 100: far inside the 1e-4 p.u. target


def test_kirchhoff_at_batch_scale(cuda_required):
    from grid2op_b200.gridmodel import GridModel
    from grid2op_b200.rollout import BatchedDoNothing
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    B = 4096
    env = BatchedDoNothing(gm, chron, B)
    worst_p = worst_q = 0.0
    for k in range(3):
        env.step_device()
        out, status, _, _ = env.fetch()
        assert (status == 0).all()
        rows = chron[env.scen, (env.t0.astype(np.int64) + k) % chron.shape[1]].astype(np.float64)
        p, q = kcl_residual(gm, out, rows)
        worst_p = max(worst_p, float(np.abs(p).max())); worst_q = max(worst_q, float(np.abs(q).max()))
    env.close()
    _dump("batch_4096x3", {"instances": B, "steps": 3, "max_abs_p_bus_MW": worst_p, "max_abs_q_bus_MVAr": worst_q,
                           "max_pu_sn100": max(worst_p, worst_q) / 100.0, "note": "float32 result records: ~1e-5 MW rounding per term"})
    assert worst_p <= 1e-2 and worst_q <= 1e-2
    assert max(worst_p, worst_q) / 100.0 <= 1e-5
