"""Planned sparse kernel, HOST build (tests/emu/sparse_emu.cpp compiles grid2op_b200/csrc/b200pf_sparse.cuh with
B200PF_EMULATE: same source as the CUDA kernel, a phase = a loop over the lanes) against the fp64 oracle: checks the
topology plans (ordering, filled pattern, factorisation schedule, factorised DC matrix) and the kernel's arithmetic on
a machine without a GPU.  The GPU run of the same kernel is checked in tests/test_engine_random_gpu.py."""
import os

import numpy as np
import pytest

from conftest import env_grid

from grid2op_b200.gridmodel import GridModel
from grid2op_b200.engine import OutputView
from oracle.c_oracle import COracle
from sparse_emu import SparseEmu
from test_c_oracle import random_cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def compare(gm, out, ref, ok):
    a, b = OutputView(gm, out[ok]), OutputView(gm, ref[ok])
    tol_mw = 1e-4 * gm.sn_mva
    for k in ("p_or", "q_or", "p_ex", "q_ex", "unit_p", "unit_q", "shunt_p", "shunt_q"):
        x, y = getattr(a, k), getattr(b, k)
        assert np.max(np.abs(x - y) - 4e-7 * np.abs(y), initial=0.0) <= tol_mw, k
    for k, vn in (("v_or", gm.line_or_vn), ("v_ex", gm.line_ex_vn), ("load_v", gm.load_vn)):
        x, y = getattr(a, k), getattr(b, k)
        assert np.max(np.abs(x - y) / vn, initial=0.0) <= 1e-4, k
    for k in ("theta_or", "theta_ex", "load_theta", "unit_theta"):
        assert np.max(np.abs(getattr(a, k) - getattr(b, k)), initial=0.0) <= 1e-3, k
    x, y = a.a_or, b.a_or
    assert np.max(np.abs(x - y) - 1e-5 * np.abs(y), initial=0.0) <= 1e-2


@pytest.mark.parametrize("name,n", [("rte_case5_example", 256), ("l2rpn_case14_sandbox", 384), ("educ_case14_storage", 128),
                                    ("l2rpn_2019", 96), ("l2rpn_neurips_2020_track1", 96), ("l2rpn_wcci_2022_dev", 24)])
@pytest.mark.parametrize("dc", [False, True])
def test_emulated_sparse_kernel_vs_oracle(name, n, dc):
    path = env_grid(name)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    topo, inj = random_cases(gm, n, seed=23)
    out, status, iters, _ = SparseEmu(gm).run(topo, inj, is_dc=dc)
    ref, rstatus, riters, _ = COracle(gm).run(topo, inj, is_dc=dc)
    assert np.array_equal(status, rstatus)                     # convergence and failure classes agree exactly
    bad = status != 0
    assert np.isnan(out[bad]).all()
    ok = ~bad
    assert ok.sum() >= n // 4
    compare(gm, out, ref, ok)
    if not dc:
        assert np.all(iters[ok] >= riters[ok]) and np.all(iters[ok] <= riters[ok] + 1)


def test_emulated_sparse_kernel_golden_fixture():
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    z = np.load(os.path.join(GOLD, "oracle_case14_steps.npz"))
    emu = SparseEmu(gm)
    out, status, iters, busv = emu.run(z["topo"], z["inj"], want_busv=True)
    assert (status == 0).all()
    compare(gm, out, z["out"], np.ones(len(out), dtype=bool))
    m = np.isfinite(z["busv"])
    assert np.array_equal(np.isfinite(busv), m)
    assert np.max(np.abs(busv[m] - z["busv"][m])) <= 1e-7
    assert np.array_equal(iters, z["iters"])
    nb, d, nnz, n_pass, n_op, n_ulev, smem, nbytes = emu.stats.tolist()
    assert nb == 14 and d == 22 and nnz < d * d // 2          # the filled Jacobian stays sparse


def test_emulated_n1_sweep_vs_oracle():
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    z = np.load(os.path.join(GOLD, "oracle_case14_steps.npz"))
    topo, inj = z["topo"][:6], z["inj"][:6]
    thl = gm.thermal_limit_a.astype(np.float32)
    out, status, iters, _, rho = SparseEmu(gm).run(topo, inj, n1_lines=gm.n_line, th_lim=thl)
    co = COracle(gm)
    pos_or, pos_ex = np.asarray(gm.line_or_pos), np.asarray(gm.line_ex_pos)
    for s in range(len(topo)):
        for l in range(gm.n_line):
            t2 = topo[s].copy()
            t2[pos_or[l]] = -1
            t2[pos_ex[l]] = -1
            ref, rs, _, _ = co.run(t2[None], inj[s][None])
            k = s * gm.n_line + l
            assert status[k] == rs[0]
            if rs[0] == 0:
                a_or = OutputView(gm, ref).a_or[0]
                assert np.allclose(rho[k], a_or / thl, rtol=2e-5, atol=1e-6)
            else:
                assert np.isnan(rho[k]).all()


@pytest.mark.parametrize("name,width", [("rte_case5_example", 32), ("l2rpn_case14_sandbox", 32), ("l2rpn_neurips_2020_track1", 32),
                                        ("l2rpn_neurips_2020_track1", 64), ("l2rpn_wcci_2022_dev", 64), ("l2rpn_wcci_2022_dev", 128)])
def test_operation_stream_is_hazard_free_and_solves_the_system(name, width):
    """Structural check of the list-scheduled operation stream (no slot of a pass touches what another slot of the pass
    writes, barrier flags on the last row of every pass) and a numeric one (the stream, run in fp64 on a random matrix with
    the plan's pattern, solves A x = b), on the reference topology, random bus splits and single-line outages."""
    from sparse_emu import validate_plan
    gm = GridModel.from_npz(os.path.join(GOLD, f"gridmodel_{name}.npz"))
    topo, _ = random_cases(gm, 24, seed=7)
    rows = [gm.default_topo()] + [topo[i] for i in range(len(topo))]
    checked = 0
    for r in rows:
        for outage in (-1, 0, gm.n_line - 1):
            rc, err = validate_plan(gm, r, outage, width)
            assert rc in (0, -1), (rc, err)          # -1: the topology has no solvable plan (islanded / no reference)
            checked += rc == 0
            rc2, err2 = validate_plan(gm, r, outage, -width)      # same with the bank-conflict-optimised layout
            assert rc2 == rc, (rc2, err2)
    assert checked >= 10


def test_emulated_protections_replay_the_recorded_rollouts():
    """Planned kernel + host cascade loop (a tripped line = a new topology = a new plan; one launch per cascade round)
    against the reference's recorded DoNothing rollouts of rte_case5_example (PandaPowerBackend with protections, 7930
    steps): same trips, same game-over step, same observed values.  CPU counterpart of tests/test_protections_gpu.py."""
    import json
    from conftest import grid2op_root
    from sparse_emu import EmuProtRollout
    root = grid2op_root()
    stat = os.path.join(root, "data", "rte_case5_example", "_statistics") if root else None
    if stat is None or not os.path.isdir(stat):
        pytest.skip("reference rollouts not available")
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_rte_case5_example.npz"))
    chron = np.load(os.path.join(GOLD, "case5_chronics.npz"))["chron"]
    meta = json.load(open(os.path.join(stat, "metadata.json")))
    keys = ["p_or", "q_or", "v_or", "a_or", "p_ex", "q_ex", "prod_p", "prod_q", "rho", "line_status", "timestep_overflow"]
    gold = {k: np.load(os.path.join(stat, f"obs_{k}.npz"))["data"] for k in keys}
    sid = np.load(os.path.join(stat, "scenario_ids.npz"))["data"].ravel().astype(int)
    rows = {s: np.flatnonzero(sid == s) for s in range(20)}
    nb_step = np.array([meta[str(s)]["nb_step"] for s in range(20)])
    B = 20
    env = EmuProtRollout(gm, chron, np.arange(B), np.zeros(B), gm.thermal_limit_a, hard=2.0, soft=1.0, max_allowed=2)
    worst = {k: 0.0 for k in keys}
    first_done = np.full(B, -1)
    n_cmp = n_casc = 0
    for k in range(int(nb_step.max())):
        n_casc += env.step(from_reset=(k == 0)) - 1
        v = OutputView(gm, env.out)
        for s in range(B):
            if env.done[s] and first_done[s] < 0:
                first_done[s] = k
            if k >= nb_step[s] - 1:
                continue
            assert env.status[s] == 0, (s, k, env.status[s])
            r = rows[s][k]
            mine = dict(p_or=v.p_or[s], q_or=v.q_or[s], v_or=v.v_or[s], a_or=v.a_or[s], p_ex=v.p_ex[s], q_ex=v.q_ex[s],
                        prod_p=v.unit_p[s], prod_q=v.unit_q[s], rho=env.rho[s])
            for name, val in mine.items():
                worst[name] = max(worst[name], float(np.max(np.abs(val.astype(np.float64) - gold[name][r]))))
            assert np.array_equal(v.a_or[s] > 0, gold["line_status"][r]) or np.array_equal(
                (np.abs(v.p_or[s]) > 0) | (v.v_or[s] > 0), gold["line_status"][r]), (s, k)
            assert np.array_equal(env.ts_over[s], gold["timestep_overflow"][r]), (s, k)
            n_cmp += 1
    for s in range(B):
        if nb_step[s] < chron.shape[1]:
            assert first_done[s] == nb_step[s] - 1, (s, first_done[s], nb_step[s])
    assert n_cmp > 7000 and n_casc > 0
    for name in ("p_or", "q_or", "p_ex", "q_ex", "prod_p", "prod_q"):
        assert worst[name] <= 1e-4, (name, worst[name])
    assert worst["v_or"] <= 2e-5 and worst["a_or"] <= 2e-3 and worst["rho"] <= 1e-6


def test_optimised_layout_gives_the_same_results_with_fewer_bank_conflicts(monkeypatch):
    """The layout optimisation only permutes where the Jacobian entries live in shared memory: results are unchanged, the
    operation stream needs fewer shared-memory wavefronts."""
    import ctypes as C
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_neurips_2020_track1.npz"))
    topo, inj = random_cases(gm, 48, seed=5)
    base = SparseEmu(gm).run(topo, inj)
    monkeypatch.setenv("SPARSE_EMU_OPTIMIZE_LAYOUT", "1")
    emu = SparseEmu(gm)
    opt = emu.run(topo, inj)
    assert np.array_equal(base[1], opt[1]) and np.array_equal(base[2], opt[2])
    ok = base[1] == 0
    assert np.allclose(base[0][ok], opt[0][ok], rtol=1e-6, atol=1e-5)
    emu.lib.sparse_emu_bank_conflicts.restype = C.c_double
    t = np.ascontiguousarray(gm.default_topo(), dtype=np.int8)
    args = (C.byref(emu.desc), t.ctypes.data_as(C.c_void_p), C.c_int(-1), C.c_int(32))
    before = emu.lib.sparse_emu_bank_conflicts(*args, C.c_int(0))
    after = emu.lib.sparse_emu_bank_conflicts(*args, C.c_int(1))
    assert 1.0 <= after < before


@pytest.mark.parametrize("name,width", [("l2rpn_case14_sandbox", 32), ("l2rpn_neurips_2020_track1", 32), ("l2rpn_wcci_2022_dev", 64)])
def test_plan_search_is_deterministic_and_never_worse(name, width):
    import ctypes as C
    gm = GridModel.from_npz(os.path.join(GOLD, f"gridmodel_{name}.npz"))
    emu = SparseEmu(gm)
    topo, _ = random_cases(gm, 6, seed=2)
    for r in [gm.default_topo()] + [topo[i] for i in range(len(topo))]:
        t = np.ascontiguousarray(r, dtype=np.int8)
        c1, c2 = C.c_uint64(0), C.c_uint64(0)
        args = (C.byref(emu.desc), t.ctypes.data_as(C.c_void_p), C.c_int(-1), C.c_int(width))
        base = emu.lib.sparse_emu_plan_rows(*args, C.c_int(0), None)
        a = emu.lib.sparse_emu_plan_rows(*args, C.c_int(1), C.byref(c1))
        b = emu.lib.sparse_emu_plan_rows(*args, C.c_int(1), C.byref(c2))
        assert a == b and c1.value == c2.value          # same plan, byte for byte
        assert (base < 0 and a < 0) or (0 < a <= base)  # rows of the operation stream


def test_plan_blobs_are_deterministic():
    """the builder is a pure function of (grid, topology record, outage, kernel shape): same input -> byte-identical blob
    (tests/emu: plan_emu_blob_hash; how builder refactorings are checked against the previous build)"""
    import ctypes as C
    import sparse_emu
    from grid2op_b200.gridmodel import GridModel
    gm = GridModel.from_npz(os.path.join(os.path.dirname(__file__), "golden", "gridmodel_l2rpn_neurips_2020_track1.npz"))
    emu = sparse_emu.SparseEmu(gm)
    emu.lib.plan_emu_blob_hash.restype = C.c_ulonglong
    t0 = np.ascontiguousarray(gm.default_topo(), dtype=np.int8)
    t1 = t0.copy(); t1[gm.line_or_pos[3]] = 2; t1[gm.line_ex_pos[5]] = 2

    def h(tv, w, T, U, seeds=0, outage=-1):
        sz = C.c_int()
        v = emu.lib.plan_emu_blob_hash(C.byref(emu.desc), tv.ctypes.data_as(C.c_void_p), outage, w, T, U, seeds, C.byref(sz))
        return int(v), sz.value

    for shape in ((64, 0, 0), (32, 0, 0), (32, 8, 1), (32, 32, 1)):
        a, b = h(t0, *shape), h(t0, *shape)
        assert a == b and a[1] > 0
        assert h(t1, *shape) != a and h(t0, *shape, outage=2) != a
    assert h(t0, 64, 0, 0, seeds=4) == h(t0, 64, 0, 0, seeds=4)
