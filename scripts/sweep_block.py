#!/usr/bin/env python
"""Tuning sweep of the block-planned kernel: (lanes per instance T, operations per lane U) x batch on the golden grid
fixtures, device-resident DoNothing stepping on the real chronics, CUDA-event timed.  The scalar planned kernel
(B200PF_BLOCK=0) is the first line of every group.    python scripts/sweep_block.py [case14|n36|wcci ...] > gpurun_out/sweep_block.json"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from grid2op_b200.gridmodel import GridModel  # noqa: E402
from grid2op_b200.rollout import BatchedDoNothing  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
CASES = {
    "case14": ("gridmodel_l2rpn_case14_sandbox.npz", "case14_sandbox_chronics.npz", (4096, 16384, 65536),
               [None, (4, 2), (2, 4), (4, 1), (4, 4), (8, 1), (8, 2), (16, 1), (16, 2), (32, 1)]),
    "n36": ("gridmodel_l2rpn_neurips_2020_track1.npz", "neurips_2020_track1_chronics.npz", (1024, 8192),
            [None, (8, 2), (4, 2), (4, 4), (8, 1), (16, 1), (16, 2), (32, 1), (32, 2)]),
    "wcci": ("gridmodel_l2rpn_wcci_2022_dev.npz", "wcci_2022_dev_chronics.npz", (1024, 2048, 8192),
             [None, (32, 1), (32, 2), (64, 1), (16, 2), (16, 1)]),
}


def rate(gm, chron, B, steps=30):
    env = BatchedDoNothing(gm, chron, B)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    env.engine.set_stream(stream.cuda_stream)
    for _ in range(3):
        env.step_device()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step_device(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    _, status, iters, _ = env.fetch()
    info = {"us_per_step_median": float(np.median(ts)) * 1e6, "us_min": float(np.min(ts)) * 1e6, "Msteps_per_s": B / float(np.median(ts)) / 1e6,
            "converged": float((status == 0).mean()), "mean_iters": float(iters.mean()), "launch": env.engine.last_launch_info(),
            "kernel": env.engine.plan_stats()["last_kernel"]}
    env.engine.set_stream(0)
    env.close()
    return info


res = {}
for case in (sys.argv[1:] or list(CASES)):
    gfile, cfile, batches, variants = CASES[case]
    gm = GridModel.from_npz(os.path.join(GOLD, gfile))
    chron = np.load(os.path.join(GOLD, cfile))["chron"]
    for var in variants:
        if var is None:
            os.environ["B200PF_BLOCK"] = "0"
        else:
            os.environ["B200PF_BLOCK"] = "1"; os.environ["B200PF_BLOCK_T"] = str(var[0]); os.environ["B200PF_BLOCK_U"] = str(var[1])
        for B in batches:
            try:
                r = rate(gm, chron, B)
            except Exception as exc:   # noqa: BLE001
                r = {"error": str(exc)}
            key = f"{case} B={B} " + ("scalar" if var is None else f"T={var[0]} U={var[1]}")
            res[key] = r
            print(key, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k != "launch"}, r.get("launch"), file=sys.stderr, flush=True)
print(json.dumps(res, indent=1))
