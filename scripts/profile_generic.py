#!/usr/bin/env python
"""One batched power flow on a golden grid fixture, for `ncu -k regex:pf_kernel -c 1`:
    python scripts/profile_generic.py l2rpn_wcci_2022_dev 296 [random_topo]"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from grid2op_b200.engine import PowerFlowEngine  # noqa: E402
from grid2op_b200.gridmodel import GridModel  # noqa: E402
from scripts.bench_configs import jitter, sub_of_pos  # noqa: E402

name, B = sys.argv[1], int(sys.argv[2])
gm = GridModel.from_npz(os.path.join(REPO, "tests", "golden", f"gridmodel_{name}.npz"))
topo = np.tile(gm.default_topo(), (B, 1))
if len(sys.argv) > 3:
    rng = np.random.default_rng(1)
    sp = sub_of_pos(gm)
    for i in range(B):
        pos = np.flatnonzero(sp == rng.integers(0, gm.n_sub))
        topo[i, pos] = rng.integers(1, 3, len(pos))
inj = jitter(gm, B)
eng = PowerFlowEngine(gm, max_batch=B)
cap = eng.max_active_buses(topo)
reps = int(os.environ.get("REPS", "1"))
eng.run(topo, inj, nb_cap=cap)
t = time.perf_counter()
for _ in range(reps):
    out, status, iters, _ = eng.run(topo, inj, nb_cap=cap)
dt = (time.perf_counter() - t) / reps
print(name, B, "s/call", dt, "inst/s", B / dt, "conv", float((status == 0).mean()), "iters", float(iters.mean()), eng.last_launch_info())
