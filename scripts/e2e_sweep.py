#!/usr/bin/env python
"""End-to-end (host buffers in / host buffers out) stepping rate of the case14 workload for several group counts and
group flags (include/b200pf.h: 1 direct out, 2 zero-copy inputs, 4 direct status).  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from grid2op_b200.gridmodel import GridModel  # noqa: E402
from grid2op_b200.rollout import BatchedDoNothing  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
B, steps = int(os.environ.get("BATCH", "4096")), 200
res = {}
for flags in (0, 2, 6, 4):
    for ng in (2, 3, 4, 6, 8):
        env = BatchedDoNothing(gm, chron, B)
        env.precollate()
        groups = env.host_groups(ng, direct_out=flags)
        n = len(groups)
        for g in range(n):
            env.group_launch(g)
        for _ in range(3):
            for g in range(n):
                env.group_wait(g); env.group_launch(g)
        for g in range(n):
            env.group_wait(g)
        t0 = time.perf_counter()
        bad = 0
        for g in range(n):
            env.group_launch(g)
        for k in range(steps):
            for g in range(n):
                out, st = env.group_wait(g)
                bad += int((st != 0).sum())
                if k + 1 < steps:
                    env.group_launch(g)
        dt = time.perf_counter() - t0
        res[f"flags={flags} groups={n}"] = B * steps / dt / 1e6
        print(f"flags={flags} groups={n}: {B * steps / dt / 1e6:.2f} M env.step/s (bad {bad})", file=sys.stderr, flush=True)
        env.close()
print(json.dumps(res, indent=1))
