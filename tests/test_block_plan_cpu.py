"""BLOCK-planned kernel, HOST build (tests/emu/sparse_emu.cpp compiles grid2op_b200/csrc/b200pf_block.cuh with
B200PF_EMULATE: same source as the CUDA kernel, the lanes of an instance as a loop, the instances of a warp one after the
other on the same element-interleaved workspace) against the fp64 oracle, and the validator of the block operation
stream (hazard-free passes, type-homogeneous rows, barrier flags, numeric solve).  GPU counterpart: test_block_kernel_gpu.py."""
import os

import numpy as np
import pytest

from conftest import env_grid

from grid2op_b200.gridmodel import GridModel
from oracle.c_oracle import COracle
from sparse_emu import BlockEmu, validate_block_plan
from test_c_oracle import random_cases
from test_sparse_plan_cpu import compare

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,n,T,U", [("rte_case5_example", 200, 4, 2), ("l2rpn_case14_sandbox", 300, 4, 2), ("l2rpn_case14_sandbox", 100, 1, 4),
                                        ("educ_case14_storage", 128, 8, 1), ("l2rpn_2019", 96, 2, 4), ("l2rpn_neurips_2020_track1", 96, 8, 2),
                                        ("l2rpn_neurips_2020_track1", 40, 16, 1), ("l2rpn_wcci_2022_dev", 24, 32, 1), ("l2rpn_wcci_2022_dev", 12, 64, 1)])
@pytest.mark.parametrize("dc", [False, True])
def test_emulated_block_kernel_vs_oracle(name, n, T, U, dc):
    path = env_grid(name)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    topo, inj = random_cases(gm, n, seed=29)
    out, status, iters, busv = BlockEmu(gm, T, U).run(topo, inj, is_dc=dc, want_busv=True)
    ref, rstatus, riters, _ = COracle(gm).run(topo, inj, is_dc=dc)
    assert np.array_equal(status, rstatus)                     # convergence and failure classes agree exactly
    bad = status != 0
    assert np.isnan(out[bad]).all()
    ok = ~bad
    assert ok.sum() >= n // 4
    compare(gm, out, ref, ok)
    if not dc:
        assert np.all(iters[ok] >= riters[ok]) and np.all(iters[ok] <= riters[ok] + 1)


def test_block_variants_are_bit_equal():
    """(T, U) only change how the same operations are packed into rows: identical results"""
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    topo, inj = random_cases(gm, 64, seed=3)
    outs = [BlockEmu(gm, T, U).run(topo, inj) for T, U in ((4, 2), (8, 1), (1, 4), (32, 1), (16, 1))]
    for o in outs[1:]:
        ok = outs[0][1] == 0
        assert np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2])
        assert np.array_equal(o[0][ok], outs[0][0][ok])


def test_block_kernel_golden_fixture_and_n1():
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    z = np.load(os.path.join(GOLD, "oracle_case14_steps.npz"))
    emu = BlockEmu(gm, 4, 2)
    out, status, iters, busv = emu.run(z["topo"], z["inj"], want_busv=True)
    assert (status == 0).all()
    compare(gm, out, z["out"], np.ones(len(out), dtype=bool))
    m = np.isfinite(z["busv"])
    assert np.array_equal(np.isfinite(busv), m)
    assert np.max(np.abs(busv[m] - z["busv"][m])) <= 1e-7
    assert np.array_equal(iters, z["iters"])
    topo, inj = z["topo"][:5], z["inj"][:5]
    thl = gm.thermal_limit_a.astype(np.float32)
    out, status, iters, _, rho = emu.run(topo, inj, n1_lines=gm.n_line, th_lim=thl)
    co = COracle(gm)
    pos_or, pos_ex = np.asarray(gm.line_or_pos), np.asarray(gm.line_ex_pos)
    from grid2op_b200.engine import OutputView
    for s in range(len(topo)):
        for l in range(gm.n_line):
            t2 = topo[s:s + 1].copy()
            t2[0, pos_or[l]] = -1; t2[0, pos_ex[l]] = -1
            ref, rst, _, _ = co.run(t2, inj[s:s + 1])
            k = s * gm.n_line + l
            assert status[k] == rst[0]
            if rst[0] == 0:
                assert np.allclose(rho[k], OutputView(gm, ref).a_or[0] / thl, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["rte_case5_example", "l2rpn_case14_sandbox", "l2rpn_neurips_2020_track1", "l2rpn_wcci_2022_dev"])
def test_block_operation_stream_is_valid(name):
    path = env_grid(name)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    topo, _ = random_cases(gm, 6, seed=41)
    big = gm.n_line > 100
    for T, U in (((32, 1), (64, 1), (16, 1)) if big else ((4, 2), (8, 1), (2, 4), (32, 1))):
        for i in range(len(topo)):
            rc, err, info = validate_block_plan(gm, topo[i], -1, T, U)
            assert rc in (0, -1), (name, T, U, i, rc, err)           # -1: the topology has no plan (islanded / no reference)
    rc, err, info = validate_block_plan(gm, gm.default_topo(), 0 if gm.n_line > 8 else -1, 32, 1)
    assert rc in (0, -1)


@pytest.mark.parametrize("name", ["l2rpn_case14_sandbox", "l2rpn_neurips_2020_track1"])
def test_three_busbars_per_substation(name):
    """``n_busbar_per_sub = 3`` (reference: Backend.can_handle_more_than_2_busbar, grid2op/Backend/backend.py:212-260; exercised through
    the backend by aaa_test_backend_interface.py:1632 test_30_n_busbar_per_sub_ok): topology plans and both planned kernels (host
    builds) on records that use busbar 3, against the fp64 oracle."""
    from sparse_emu import SparseEmu
    from test_c_oracle import _sub_of
    path = env_grid(name)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path, n_busbar=3)
    rng = np.random.default_rng(7)
    n = 120
    topo = np.tile(gm.default_topo(), (n, 1))
    inj = np.tile(gm.default_inj(), (n, 1))
    for i in range(1, n):
        for s in rng.choice(gm.n_sub, size=rng.integers(0, 3), replace=False):
            for p in range(gm.dim_topo):
                if _sub_of(gm, p) == s and topo[i, p] > 0:
                    topo[i, p] = rng.integers(1, 4)
    assert (topo == 3).any()
    ref, rstatus, _, _ = COracle(gm).run(topo, inj)
    assert (rstatus == 0).sum() >= n // 4 and (rstatus != 0).any()
    for emu in (SparseEmu(gm), BlockEmu(gm, 8, 1)):
        out, status, _, _ = emu.run(topo, inj)
        assert np.array_equal(status, rstatus)
        compare(gm, out, ref, status == 0)


@pytest.mark.parametrize("name", ["educ_case14_redisp", "l2rpn_case14_sandbox_diff_grid", "l2rpn_icaps_2021", "l2rpn_idf_2023",
                                  "l2rpn_neurips_2020_track2", "l2rpn_wcci_2020", "rte_case118_example", "rte_case14_opponent",
                                  "rte_case14_realistic", "rte_case14_redisp", "rte_case14_test"])
def test_every_other_bundled_grid_through_the_planned_kernels(name):
    """the remaining grid files the reference bundles (other line / transformer parameters, 118-substation variants): topology plans +
    both planned kernels (host builds) against the fp64 oracle on random splits / outages / injections"""
    from sparse_emu import SparseEmu
    from conftest import grid2op_root
    root = grid2op_root()
    path = None
    if root is not None:
        base = os.path.join(root, "data", name)
        cands = [os.path.join(base, "grid.json")] + ([os.path.join(base, d, "grid.json") for d in sorted(os.listdir(base))] if os.path.isdir(base) else [])
        path = next((p for p in cands if os.path.exists(p)), None)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    n = 12 if gm.n_sub > 50 else 32
    topo, inj = random_cases(gm, n, seed=5)
    ref, rstatus, _, _ = COracle(gm).run(topo, inj)
    assert (rstatus == 0).sum() >= n // 4
    for emu in (SparseEmu(gm), BlockEmu(gm, 8, 1) if gm.n_line <= 32 else BlockEmu(gm, 32, 1)):
        out, status, _, _ = emu.run(topo, inj)
        assert np.array_equal(status, rstatus)
        compare(gm, out, ref, status == 0)
