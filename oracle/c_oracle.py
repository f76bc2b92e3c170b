"""ORACLE (test infrastructure / CPU baseline): ctypes front-end of oracle/pf_oracle.c.

``COracle.run`` has the signature of ``grid2op_b200.engine.PowerFlowEngine.run`` so tests and the
CPU-baseline legs of bench.py can swap it in.  Build: ``python oracle/build.py`` (gcc -O3 -fopenmp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpf_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "pf_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", LIB, src, "-lm"]
        subprocess.check_call(cmd)
    return LIB


class COracle:
    def __init__(self, gm, nthreads: int = 0):
        from grid2op_b200.engine import _GridDesc   # struct layout of include/b200pf.h (types only)
        build()
        self.gm = gm
        self.lib = C.CDLL(LIB)
        self.lib.pf_oracle_run.restype = C.c_int
        self.lib.pf_oracle_max_threads.restype = C.c_int
        self.nthreads = int(nthreads)
        self._keep = []
        d = _GridDesc()
        d.abi_version = 2
        d.n_sub, d.n_busbar = gm.n_sub, gm.n_busbar
        d.n_line, d.n_gen, d.n_hidden, d.n_load = gm.n_line, gm.n_gen, gm.n_hidden, gm.n_load
        d.n_storage, d.n_shunt, d.dim_topo = gm.n_storage, gm.n_shunt, gm.dim_topo
        d.sn_mva = gm.sn_mva

        def put(name, arr, dtype):
            a = np.ascontiguousarray(arr, dtype=dtype)
            if a.size == 0:
                a = np.zeros(1, dtype=dtype)
            self._keep.append(a)
            setattr(d, name, a.ctypes.data)

        i32, f64, f32 = np.int32, np.float64, np.float32
        for nm, arr, dt in (
                ("line_or_sub", gm.line_or_sub, i32), ("line_ex_sub", gm.line_ex_sub, i32), ("line_or_pos", gm.line_or_pos, i32),
                ("line_ex_pos", gm.line_ex_pos, i32), ("line_y", gm.line_y, f64), ("line_bdc", gm.line_bdc, f64),
                ("line_pshift", gm.line_pshift, f64), ("line_or_vn", gm.line_or_vn, f32), ("line_ex_vn", gm.line_ex_vn, f32),
                ("unit_sub", gm.unit_sub, i32), ("unit_pos", gm.unit_pos, i32), ("unit_is_ref", gm.unit_is_ref, i32),
                ("unit_qmin", gm.unit_qmin, f64), ("unit_qmax", gm.unit_qmax, f64), ("unit_vn", gm.unit_vn, f32),
                ("load_sub", gm.load_sub, i32), ("load_pos", gm.load_pos, i32), ("load_vn", gm.load_vn, f32),
                ("storage_sub", gm.storage_sub, i32), ("storage_pos", gm.storage_pos, i32), ("storage_vn", gm.storage_vn, f32),
                ("storage_q", gm.storage_q, f64), ("shunt_sub", gm.shunt_sub, i32), ("shunt_vn", gm.shunt_vn, f32),
                ("shunt_vratio", gm.shunt_vratio, f64)):
            put(nm, arr, dt)
        self.desc = d

    @property
    def max_threads(self) -> int:
        return int(self.lib.pf_oracle_max_threads())

    def run(self, topo, inj, is_dc=False, max_iter=10, tol_mva=1e-8, nb_cap=0, want_busv=False):
        gm = self.gm
        topo = np.ascontiguousarray(topo, dtype=np.int8).reshape(-1, gm.n_topo_in)
        inj = np.ascontiguousarray(inj, dtype=np.float64).reshape(-1, gm.n_inj)
        B = topo.shape[0]
        out = np.empty((B, gm.n_out), dtype=np.float32)
        status = np.empty(B, dtype=np.int32)
        iters = np.empty(B, dtype=np.int32)
        vp = C.c_void_p
        self.lib.pf_oracle_run(C.byref(self.desc), C.c_int(B), topo.ctypes.data_as(vp), inj.ctypes.data_as(vp),
                               C.c_int(int(bool(is_dc))), C.c_int(int(max_iter)), C.c_double(float(tol_mva)),
                               out.ctypes.data_as(vp), status.ctypes.data_as(vp), iters.ctypes.data_as(vp),
                               C.c_int(self.nthreads))
        return out, status, iters, None


if __name__ == "__main__":
    print(build(force=True))
