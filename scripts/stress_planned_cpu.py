"""TEST TOOL (CPU only): large randomised comparison of the planned sparse kernel's HOST build (tests/emu) with the fp64
pivoting oracle (oracle/pf_oracle.c) — status classes, iteration counts, values — to look for break-downs of the
non-pivoting elimination.  SPARSE_EMU_OPTIMIZE_LAYOUT=2 python scripts/stress_planned_cpu.py  (searched plans)."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from conftest import env_grid
from grid2op_b200.gridmodel import GridModel
from oracle.c_oracle import COracle
from test_c_oracle import random_cases
from sparse_emu import SparseEmu
tot = mism = 0
for name, n in [("rte_case5_example", 3000), ("l2rpn_case14_sandbox", 3000), ("educ_case14_storage", 2000), ("l2rpn_2019", 2000),
                ("l2rpn_neurips_2020_track1", 1500), ("l2rpn_wcci_2022_dev", 300)]:
    gm = GridModel(env_grid(name))
    emu, co = SparseEmu(gm), COracle(gm)
    for seed in (101, 202):
        topo, inj = random_cases(gm, n, seed=seed)
        out, st, it, _ = emu.run(topo, inj)
        ref, rs, ri, _ = co.run(topo, inj)
        bad = st != rs
        ok = (st == 0) & (rs == 0)
        di = it[ok] - ri[ok]
        err = np.max(np.abs(out[ok] - ref[ok]) / (1 + np.abs(ref[ok]))) if ok.any() else 0
        tot += n; mism += int(bad.sum())
        print(f"{name:28s} seed {seed}: n={n} converged {int(ok.sum())} status mismatches {int(bad.sum())} iters diff [{di.min()},{di.max()}] max rel err {err:.2e}", flush=True)
        if bad.any():
            idx = np.flatnonzero(bad)[:5]
            print("    mismatch", idx, st[idx], rs[idx], it[idx], ri[idx])
print("total", tot, "mismatches", mism)
