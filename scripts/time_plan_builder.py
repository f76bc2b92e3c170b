"""Host-side cost of a topology plan (csrc/b200pf_plan.hpp), one thread, through the emulation build of the kernel sources
(tests/emu: no GPU needed).  Topologies: the random walk of BASELINE configs[2] (one random substation re-assignment per step).

    python scripts/time_plan_builder.py [grid npz] [n]
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))


def random_walk_topologies(gm, n, rng, restart_every=4):
    sub_of = np.full(gm.dim_topo, -1, dtype=np.int64)
    sub_of[gm.line_or_pos] = gm.line_or_sub; sub_of[gm.line_ex_pos] = gm.line_ex_sub
    sub_of[gm.gen_pos] = gm.gen_sub; sub_of[gm.load_pos] = gm.load_sub
    if gm.n_storage:
        sub_of[gm.storage_pos] = gm.storage_sub
    rows = np.empty((n, gm.n_topo_in), dtype=np.int8)
    cur = gm.default_topo().copy()
    for k in range(n):
        if k % restart_every == 0:
            cur = gm.default_topo().copy()
        s = int(rng.integers(0, gm.n_sub))
        pos = np.flatnonzero(sub_of == s)
        cur[pos] = rng.integers(1, 3, len(pos))
        rows[k] = cur
    return rows


def main():
    from grid2op_b200.gridmodel import GridModel
    import sparse_emu
    gold = os.path.join(os.path.dirname(HERE), "tests", "golden")
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(gold, "gridmodel_l2rpn_neurips_2020_track1.npz")
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    gm = GridModel.from_npz(path)
    emu = sparse_emu.SparseEmu(gm)
    rows = random_walk_topologies(gm, n, np.random.default_rng(0))
    emu.lib.plan_emu_build_many.restype = C.c_int
    for name, w, T, U in (("scalar, 64 slots per row", 64, 0, 0), ("scalar, 32 slots per row", 32, 0, 0), ("block 32x1", 32, 32, 1), ("block 8x1", 32, 8, 1)):
        sec, nbytes, bad = C.c_double(), C.c_longlong(), C.c_int()
        emu.lib.plan_emu_build_many(C.byref(emu.desc), n, rows.ctypes.data_as(C.c_void_p), w, T, U, C.byref(sec), C.byref(nbytes), C.byref(bad))
        ok = max(1, n - bad.value)
        print(f"{name:28s} {1e6 * sec.value / ok:8.1f} us / plan   {nbytes.value / ok / 1024:6.1f} KB / plan   ({bad.value} of {n} topologies with isolated elements)")


if __name__ == "__main__":
    main()
