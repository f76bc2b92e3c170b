#!/usr/bin/env python
"""Single-environment drop-in latency: unmodified grid2op Environment + B200Backend, one instance, DoNothing agent,
NO_OVERFLOW_DISCONNECTION — the loop and the per-phase split of the reference's own profiling script
(_profiling/average_time_in_step_no_redisp.py:35-83: phases 1-2 = _BackendAction creation, 3 = apply_action, 4 = powerflow,
5 = observation extraction), plus obs.simulate() and env.reset().  For comparison the reference documents 17.3 ms
(PandaPowerBackend) and 1.56 ms (lightsim2grid) per simulate / step on l2rpn_case14_sandbox-sized grids
(docs/user/action.rst:315-321).  This is the LATENCY view of the drop-in; throughput is bench.py's batched path.

    python scripts/single_env_latency.py > gpurun_out/single_env_latency.json"""
import json
import os
import sys
import time
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from grid2op_b200.backend import B200Backend  # noqa: E402  (locates the grid2op install)
import grid2op  # noqa: E402
from grid2op.Parameters import Parameters  # noqa: E402

NB_TS = int(os.environ.get("NB_TS", "500"))
res = {}
for env_nm in ["l2rpn_case14_sandbox", "l2rpn_neurips_2020_track1", "l2rpn_wcci_2022_dev"]:
    p = Parameters()
    p.NO_OVERFLOW_DISCONNECTION = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            env = grid2op.make(env_nm, test=True, param=p, backend=B200Backend(), opponent_init_budget=0., opponent_budget_per_ts=0.)
        except TypeError:
            env = grid2op.make(env_nm, test=True, param=p, backend=B200Backend())
    obs = env.reset()
    for _ in range(10):
        env.step(env.action_space())
    env._time_create_bk_act = 0.; env._time_apply_act = 0.; env._time_powerflow = 0.; env._time_extract_obs = 0.; env._time_step = 0.
    t0 = time.perf_counter()
    n = 0
    for i in range(NB_TS):
        obs, reward, done, info = env.step(env.action_space())
        n += 1
        if done:
            break
    wall = time.perf_counter() - t0
    r = {"steps": n, "ms_per_step_wall": 1e3 * wall / n, "ms_total_env_timer": 1e3 * env._time_step / n,
         "ms_1_2_backend_action": 1e3 * env._time_create_bk_act / n, "ms_3_apply_action": 1e3 * (env._time_apply_act - env._time_create_bk_act) / n,
         "ms_4_powerflow": 1e3 * env._time_powerflow / n, "ms_5_extract_obs": 1e3 * env._time_extract_obs / n}
    # the backend call alone (apply nothing, runpf): one C-ABI call = plan lookup + H2D + one launch + D2H
    bk = env.backend
    ts = []
    for _ in range(200):
        t = time.perf_counter(); bk.runpf(is_dc=False); ts.append(time.perf_counter() - t)
    r["ms_runpf_alone_median"] = 1e3 * float(np.median(ts)); r["ms_runpf_alone_min"] = 1e3 * float(np.min(ts))
    # obs.simulate (Backend.copy + forecast injections + one power flow)
    obs = env.reset()
    ts = []
    for _ in range(100):
        t = time.perf_counter(); obs.simulate(env.action_space()); ts.append(time.perf_counter() - t)
    r["ms_simulate_median"] = 1e3 * float(np.median(ts))
    ts = []
    for _ in range(5):
        t = time.perf_counter(); env.reset(); ts.append(time.perf_counter() - t)
    r["ms_reset_median"] = 1e3 * float(np.median(ts))
    res[env_nm] = r
    print(env_nm, {k: round(v, 4) if isinstance(v, float) else v for k, v in r.items()}, file=sys.stderr, flush=True)
    env.close()
res["reference_doc"] = {"pandapower_ms_per_powerflow": 17.3, "lightsim2grid_ms_per_powerflow": 1.56, "source": "docs/user/action.rst:315-321 (reference's own numbers, other hardware)"}
print(json.dumps(res, indent=1))
