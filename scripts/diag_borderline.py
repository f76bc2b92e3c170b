#!/usr/bin/env python
"""Diagnostics of the one 118-substation state (stress chunk seed 1002, instance 3253) on which the pivoting re-solve and the
fp64 oracle disagree about convergence: status / iterations of every kernel family on it."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from conftest import env_grid  # noqa: E402
from grid2op_b200.engine import PowerFlowEngine  # noqa: E402
from grid2op_b200.gridmodel import GridModel  # noqa: E402
from oracle.c_oracle import COracle  # noqa: E402
from test_redo_gpu import fast_random_cases  # noqa: E402

gm = GridModel(env_grid("l2rpn_wcci_2022_dev"))
topo, inj = fast_random_cases(gm, 4096, seed=1002)
idx = np.array([3253, 1944, 0, 1, 2, 3, 4, 5])
t, j = topo[idx], inj[idx]
ref, rs, ri, _ = COracle(gm).run(t, j)
print("oracle        status", rs.tolist(), "iters", ri.tolist())
for it in (10, 20):
    _, s2, i2, _ = COracle(gm).run(t, j, max_iter=it)
    print(f"oracle max_iter={it}", s2.tolist(), i2.tolist())
for name, pol, env in (("planned+redo", 2, {}), ("planned, no redo", 2, {"B200PF_NO_REDO": "1"}), ("pivoting fp32 (policy 1)", 1, {})):
    for k, v in env.items():
        os.environ[k] = v
    eng = PowerFlowEngine(gm, max_batch=len(idx))
    eng.set_kernel_policy(pol)
    for mi in (10, 20):
        out, st, its, _ = eng.run(t, j, max_iter=mi)
        print(f"{name:26s} max_iter={mi} status", st.tolist(), "iters", its.tolist(), eng.plan_stats()["last_kernel"])
    eng.close()
    for k in env:
        os.environ.pop(k)
