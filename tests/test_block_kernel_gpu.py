"""BLOCK-planned kernel (csrc/b200pf_block.cuh) on the GPU: every compiled (lanes per instance, operations per lane) variant
against the fp64 oracle on random states of four grids — each instance brings its own topology, so the instances that share
a warp run different plans and different iteration counts (divergence inside a warp is part of the design) — and against
each other: the variants only pack the same operations differently, so their results must be BIT-EQUAL.  The scalar
planned kernel (B200PF_BLOCK=0) is kept under test through the same cases."""
import os

import numpy as np
import pytest

from conftest import env_grid

from grid2op_b200.gridmodel import GridModel
from oracle.c_oracle import COracle
from test_engine_random_gpu import _compare
from test_redo_gpu import fast_random_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

VARIANTS = [(4, 2), (4, 1), (8, 1), (16, 1), (32, 1), (64, 1)]       # the compiled (lanes, operations) variants


class _env:
    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _engine(gm, n, T=None, U=None, block=1):
    from grid2op_b200.engine import PowerFlowEngine
    kw = {"B200PF_BLOCK": block}
    if T is not None:
        kw.update(B200PF_BLOCK_T=T, B200PF_BLOCK_U=U)
    with _env(**kw):
        eng = PowerFlowEngine(gm, max_batch=n)
    eng.set_kernel_policy(2)
    return eng


@pytest.mark.parametrize("name,n", [("rte_case5_example", 300), ("l2rpn_case14_sandbox", 1000), ("l2rpn_neurips_2020_track1", 300),
                                    ("l2rpn_wcci_2022_dev", 100)])
@pytest.mark.parametrize("dc", [False, True])
def test_block_variants_vs_oracle_and_each_other(cuda_required, name, n, dc):
    path = env_grid(name)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    topo, inj = fast_random_cases(gm, n, seed=77)
    ref, rstatus, riters, _ = COracle(gm).run(topo, inj, is_dc=dc)
    first = None
    n_run = 0
    for T, U in VARIANTS:
        eng = _engine(gm, n, T, U)
        try:
            out, status, iters, busv = eng.run(topo, inj, is_dc=dc, want_busv=True)
        except RuntimeError:
            eng.close()
            raise
        info = eng.last_launch_info()
        kern = eng.plan_stats()["last_kernel"]
        eng.close()
        if kern != "planned_block":            # plans of this (T, U) do not fit the format for this grid: pivoting kernels ran
            continue
        assert info["threads_per_instance"] == T
        n_run += 1
        assert np.array_equal(status, rstatus), (T, U)
        ok = status == 0
        _compare(gm, out, ref, ok)
        if not dc:
            d = iters[ok] - riters[ok]
            assert d.min() >= -1 and d.max() <= 1, (T, U, d.min(), d.max())
        if first is None:
            first = (out, iters, busv)
        else:
            assert np.array_equal(out[ok], first[0][ok]), (T, U)
            assert np.array_equal(iters, first[1]), (T, U)
            assert np.array_equal(busv[ok], first[2][ok], equal_nan=True), (T, U)
    assert n_run >= 3
    # the scalar planned kernel stays a supported path
    eng = _engine(gm, n, block=0)
    out, status, iters, _ = eng.run(topo, inj, is_dc=dc)
    assert eng.plan_stats()["last_kernel"] == "planned_sparse"
    eng.close()
    assert np.array_equal(status, rstatus)
    _compare(gm, out, ref, status == 0)


def test_block_kernel_series_rows_n1_and_shared_plan_batches(cuda_required):
    """a batch that shares ONE plan (the DoNothing rollout of bench.py) with a size that is not a multiple of the instances
    per warp; series, rows and N-1 entry points; scalar planned kernel as cross-check (tolerance: both fp32 factorisations)"""
    from grid2op_b200.rollout import BatchedDoNothing
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    B = 1003
    with _env(B200PF_BLOCK=0):
        ref = BatchedDoNothing(gm, chron, B)           # scalar planned kernel, device-resident stepping
        ref_h = BatchedDoNothing(gm, chron, B)         # ... and a second one to pair with the host-path steps
    env = BatchedDoNothing(gm, chron, B)
    for k in range(4):
        ref.step_device(); env.step_device()
        o1, s1, i1, r1 = ref.fetch()
        o2, s2, i2, r2 = env.fetch()
        assert env.engine.plan_stats()["last_kernel"] == "planned_block" and ref.engine.plan_stats()["last_kernel"] == "planned_sparse"
        assert (s1 == 0).all() and (s2 == 0).all() and np.abs(i1 - i2).max() <= 1
        _compare(gm, o2, o1, s1 == 0)
        assert np.allclose(r1, r2, rtol=2e-5, atol=1e-6)
        o3, s3 = env.step_host()                       # (the host path keeps its own row counter: k-th call = k-th row)
        ref_h.step_device()
        o1, s1, _, _ = ref_h.fetch()
        _compare(gm, o3.copy(), o1, s1 == 0)
    ref_h.close()
    z = np.load(os.path.join(GOLD, "oracle_case14_steps.npz"))
    ref.close(); env.close()
    from grid2op_b200.engine import PowerFlowEngine
    with _env(B200PF_BLOCK=0):
        e0 = PowerFlowEngine(gm, max_batch=9 * gm.n_line)
    e1 = PowerFlowEngine(gm, max_batch=9 * gm.n_line)
    rho0, st0 = e0.n1_sweep(z["topo"][:9], z["inj"][:9])
    rho1, st1 = e1.n1_sweep(z["topo"][:9], z["inj"][:9])
    assert e1.plan_stats()["last_kernel"] == "planned_block"
    assert np.array_equal(st0, st1)
    ok = st0 == 0
    assert np.allclose(rho0[ok], rho1[ok], rtol=2e-5, atol=1e-6)
    e0.close(); e1.close()


def test_lockstep_warps_equal_free_running_warps(cuda_required):
    """Shared-plan launches run the block kernel with lockstep warps (full-mask barriers); the free-running variant
    (B200PF_BLOCK_UNI=0) must give bit-identical records — also for instances that fail inside a warp (test knob) while their
    warp mates go on, and with iteration counts that differ inside a warp."""
    from grid2op_b200.rollout import BatchedDoNothing
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"].copy()
    chron[1, :, :2 * gm.n_load] *= 3.5                    # scenario 1: 5 Newton iterations instead of 4; scenario 2: beyond collapse
    chron[2, :, :2 * gm.n_load] *= 4.5                    # (diverges after 10) — all three inside every warp (instance i -> scenario i mod 3)
    B = 1024
    with _env(B200PF_BLOCK_UNI=0):
        free = BatchedDoNothing(gm, chron, B)
    lock = BatchedDoNothing(gm, chron, B)
    n_counts = 0
    for knob in (0, 3):
        free.engine.set_debug(knob, redo_enabled=False); lock.engine.set_debug(knob, redo_enabled=False)
        for _ in range(3):
            free.step_device(); lock.step_device()
            o1, s1, i1, r1 = free.fetch()
            o2, s2, i2, r2 = lock.fetch()
            assert lock.engine.plan_stats()["last_kernel"] == "planned_block"
            assert np.array_equal(s1, s2) and np.array_equal(i1, i2)
            n_counts = max(n_counts, len(np.unique(i1[s1 == 0])))
            assert np.array_equal(o1, o2, equal_nan=True) and np.array_equal(r1, r2, equal_nan=True)
            assert (s1 == 1).any() and (i1[s1 == 1] == 10).any() and (s1 == 0).any()          # genuinely diverging instances among them
            if knob:
                assert (s1[::3] == 1).all()
    assert n_counts >= 2, "the case should mix iteration counts inside warps"
    free.close(); lock.close()
