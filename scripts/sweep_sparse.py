#!/usr/bin/env python
"""Tuning sweep of the planned sparse kernel (threads per instance x resident CTAs per SM) on the golden grid fixtures,
device-resident DoNothing stepping.  Usage: python scripts/sweep_sparse.py > gpurun_out/sweep.json"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from grid2op_b200.gridmodel import GridModel  # noqa: E402
from scripts.bench_configs import series_rate  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
res = {}
for name, batches, combos in (
        ("l2rpn_wcci_2022_dev", (1024, 2048, 8192), ((64, 0), (64, 7), (32, 0), (128, 0))),
        ("l2rpn_neurips_2020_track1", (1024, 8192), ((32, 0), (64, 0))),
        ("l2rpn_case14_sandbox", (4096, 16384), ((32, 0),))):
    gm = GridModel.from_npz(os.path.join(GOLD, f"gridmodel_{name}.npz"))
    for T, cap in combos:
        os.environ["B200PF_SPARSE_T"] = str(T)
        os.environ["B200PF_SPARSE_CTAS"] = str(cap)
        for B in batches:
            r = series_rate(gm, B, 2, steps=30)
            res[f"{name} B={B} T={T} cap={cap}"] = {"Msteps_per_s": r["env_step_per_s"] / 1e6, "us_per_step": r["seconds_per_step"] * 1e6,
                                                   "grid": r["launch"]["grid"], "smem": r["launch"]["smem_bytes"]}
            print(f"{name} B={B} T={T} cap={cap}: {r['env_step_per_s'] / 1e6:.3f} M/s {r['seconds_per_step'] * 1e6:.1f} us grid {r['launch']['grid']}", file=sys.stderr, flush=True)
print(json.dumps(res, indent=1))
