"""The warp-per-instance kernel (b200pf_small.cuh) and the generic group kernel (b200pf_kernel.cuh) are two
implementations of the same solve: they must agree with each other and with the oracle."""
import os

import numpy as np
import pytest

from conftest import env_grid

from grid2op_b200.gridmodel import GridModel
from oracle.c_oracle import COracle
from test_c_oracle import random_cases
from test_engine_random_gpu import _compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["rte_case5_example", "l2rpn_case14_sandbox", "educ_case14_storage", "l2rpn_2019"])
@pytest.mark.parametrize("dc", [False, True])
def test_small_vs_generic_vs_oracle(cuda_required, name, dc):
    from grid2op_b200.engine import PowerFlowEngine
    path = env_grid(name)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    n = 384
    topo, inj = random_cases(gm, n, seed=23)
    eng = PowerFlowEngine(gm, max_batch=n)
    eng.set_kernel_policy(1)                 # the two pivoting kernels first
    cap = eng.max_active_buses(topo)
    assert 0 < cap <= gm.n_slot
    if cap > 17:   # keep only the instances the small kernel is allowed to take
        keep = np.array([eng.max_active_buses(topo[i:i + 1]) <= 17 for i in range(n)])
        topo, inj = topo[keep], inj[keep]
        cap = eng.max_active_buses(topo)
    o_small, s_small, i_small, bv_small = eng.run(topo, inj, is_dc=dc, nb_cap=cap, want_busv=True)
    assert eng.last_launch_info()["block"] == 32
    o_gen, s_gen, i_gen, bv_gen = eng.run(topo, inj, is_dc=dc, nb_cap=0, want_busv=True)
    assert np.array_equal(s_small, s_gen)
    ok = s_small == 0
    assert ok.sum() >= len(ok) // 4
    assert np.isnan(o_small[~ok]).all()
    # same arithmetic plan, different summation order in places: agreement far below the parity tolerance
    assert np.max(np.abs(bv_small[ok][np.isfinite(bv_gen[ok])] - bv_gen[ok][np.isfinite(bv_gen[ok])])) <= 1e-9
    assert np.array_equal(np.isnan(bv_small), np.isnan(bv_gen))
    ref, rs, ri, _ = COracle(gm).run(topo, inj, is_dc=dc)
    assert np.array_equal(s_small == 0, rs == 0) and np.array_equal(s_small[~ok], rs[~ok])
    _compare(gm, o_small, ref, ok)
    _compare(gm, o_gen, ref, ok)
    if not dc:
        assert np.all(i_small[ok] >= ri[ok]) and np.all(i_small[ok] <= ri[ok] + 1)
    # third implementation: the planned sparse kernel (no pivoting, static schedule per topology)
    eng.set_kernel_policy(2)
    o_pl, s_pl, i_pl, bv_pl = eng.run(topo, inj, is_dc=dc, want_busv=True)
    assert eng.plan_stats()["last_kernel"].startswith("planned")
    assert np.array_equal(s_pl, s_small)
    assert np.array_equal(np.isnan(bv_pl), np.isnan(bv_gen))
    assert np.max(np.abs(bv_pl[ok][np.isfinite(bv_gen[ok])] - bv_gen[ok][np.isfinite(bv_gen[ok])])) <= 1e-9
    _compare(gm, o_pl, ref, ok)
    eng.close()
