"""GPU parity on perturbed states: random bus splits, line outages and injection changes (converging and
diverging), CUDA engine through the C ABI vs the fp64 oracle (C restatement, itself checked against the
numpy oracle in tests/test_c_oracle.py).  Tolerance: 1e-4 p.u. (north_star) on flows and voltages."""
import os

import numpy as np
import pytest

from conftest import env_grid

from grid2op_b200.gridmodel import GridModel
from oracle.c_oracle import COracle
from test_c_oracle import random_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _compare(gm, out, ref, ok, theta_tol=1e-3, theta_vmin=0.0):
    """theta_vmin > 0: angles are compared only where the voltage is above theta_vmin p.u. (the angle of a collapsed bus, |V| ~ 0,
    is not determined by the power-flow equations)"""
    from grid2op_b200.engine import OutputView
    a, b = OutputView(gm, out[ok]), OutputView(gm, ref[ok])
    tol_mw = 1e-4 * gm.sn_mva
    for k in ("p_or", "q_or", "p_ex", "q_ex", "unit_p", "unit_q", "shunt_p", "shunt_q"):
        x, y = getattr(a, k), getattr(b, k)
        assert np.max(np.abs(x - y) - 4e-7 * np.abs(y), initial=0.0) <= tol_mw, k          # + float32 record rounding
    for k, vn in (("v_or", gm.line_or_vn), ("v_ex", gm.line_ex_vn), ("load_v", gm.load_vn)):
        x, y = getattr(a, k), getattr(b, k)
        assert np.max(np.abs(x - y) / vn, initial=0.0) <= 1e-4, k                            # p.u.
    for k, kv, vn in (("theta_or", "v_or", gm.line_or_vn), ("theta_ex", "v_ex", gm.line_ex_vn), ("load_theta", "load_v", gm.load_vn),
                      ("unit_theta", "unit_v", gm.unit_vn)):
        d = np.abs(getattr(a, k) - getattr(b, k))
        d = np.minimum(d, np.abs(d - 360.0))                                                 # -180 == +180
        if theta_vmin > 0:
            d = np.where(getattr(b, kv) / vn > theta_vmin, d, 0.0)
        assert np.max(d, initial=0.0) <= theta_tol, k                                        # degrees
    x, y = a.a_or, b.a_or
    d = np.abs(x - y) - 1e-5 * np.abs(y)
    if theta_vmin > 0:                                     # (a = S / (sqrt(3) V): same remark for a collapsed bus)
        d = np.where(b.v_or / gm.line_or_vn > theta_vmin, d, 0.0)
    assert np.max(d, initial=0.0) <= 1e-2


@pytest.mark.parametrize("name,n", [("rte_case5_example", 256), ("l2rpn_case14_sandbox", 512), ("educ_case14_storage", 256),
                                    ("l2rpn_2019", 128), ("l2rpn_neurips_2020_track1", 128), ("l2rpn_wcci_2022_dev", 16)])
@pytest.mark.parametrize("dc", [False, True])
@pytest.mark.parametrize("policy", [1, 2])       # 1: pivoting kernels (warp / CTA per instance), 2: planned sparse kernel
def test_random_states(cuda_required, name, n, dc, policy):
    from grid2op_b200.engine import PowerFlowEngine
    path = env_grid(name)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    topo, inj = random_cases(gm, n, seed=11)
    eng = PowerFlowEngine(gm, max_batch=n)
    eng.set_kernel_policy(policy)
    out, status, iters, _ = eng.run(topo, inj, is_dc=dc)
    assert (eng.plan_stats()["last_kernel"].startswith("planned")) == (policy == 2)
    ref, rstatus, riters, _ = COracle(gm).run(topo, inj, is_dc=dc)
    # integers: convergence / failure class must agree exactly
    assert np.array_equal(status == 0, rstatus == 0)
    bad = status != 0
    assert np.array_equal(status[bad], rstatus[bad])
    assert np.isnan(out[bad]).all()
    ok = ~bad
    assert ok.sum() >= n // 4
    _compare(gm, out, ref, ok)
    if not dc:
        # fp32 Jacobian + fp64 residual: same iteration count as the fp64 Newton, at most one more
        assert np.all(iters[ok] >= riters[ok]) and np.all(iters[ok] <= riters[ok] + 1)
    eng.close()


@pytest.mark.parametrize("policy", [1, 2])
def test_golden_fixture_case14(cuda_required, policy):
    """Committed oracle fixture (tests/golden/oracle_case14_steps.npz, made by make_golden.py): runs without
    any reference file, e.g. on a GPU box that only has this repository."""
    from grid2op_b200.engine import PowerFlowEngine
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    z = np.load(os.path.join(GOLD, "oracle_case14_steps.npz"))
    eng = PowerFlowEngine(gm, max_batch=len(z["topo"]))
    eng.set_kernel_policy(policy)
    out, status, iters, busv = eng.run(z["topo"], z["inj"], want_busv=True)
    assert (status == 0).all()
    assert (eng.plan_stats()["last_kernel"].startswith("planned")) == (policy == 2)
    _compare(gm, out, z["out"], np.ones(len(out), dtype=bool))
    m = np.isfinite(z["busv"])
    assert np.max(np.abs(busv[m] - z["busv"][m])) <= 1e-7          # p.u. / rad, fp64 state
    assert np.array_equal(iters, z["iters"])
    eng.close()


@pytest.mark.parametrize("policy", [1, 2])
def test_series_mode_matches_explicit_records(cuda_required, policy):
    """Device-resident chronics stepping == the same rows fed through the host-record entry point."""
    from grid2op_b200.engine import PowerFlowEngine
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    B = 300
    scen = (np.arange(B) % 3).astype(np.int32)
    t0 = ((np.arange(B) * 37) % chron.shape[1]).astype(np.int32)
    eng = PowerFlowEngine(gm, max_batch=B)
    eng.set_kernel_policy(policy)
    eng.series_bind(chron, scen, t0, gm.default_inj(), gm.thermal_limit_a)
    eng.series_set_topo(np.tile(gm.default_topo(), (B, 1)))
    sl = gm.inj_slices()
    nl, ng = gm.n_load, gm.n_gen
    for step in range(3):
        eng.series_step(nb_cap=gm.n_sub)
        assert (eng.plan_stats()["last_kernel"].startswith("planned")) == (policy == 2)
        out, status, iters, rho = eng.series_fetch()
        rows = chron[scen, (t0 + step) % chron.shape[1]]
        inj = np.tile(gm.default_inj(), (B, 1))
        inj[:, sl["load_p"]] = rows[:, :nl]
        inj[:, sl["load_q"]] = rows[:, nl:2 * nl]
        inj[:, sl["gen_p"]] = rows[:, 2 * nl:2 * nl + ng]
        inj[:, sl["gen_vm"]] = (rows[:, 2 * nl + ng:] / gm.prod_pu_to_kv[None, :]).astype(np.float32)
        out2, status2, iters2, _ = eng.run(np.tile(gm.default_topo(), (B, 1)), inj)
        assert (status == 0).all() and (status2 == 0).all()
        assert np.array_equal(out, out2) and np.array_equal(iters, iters2)       # bit-exact: same kernel, same inputs
        v = eng.view(out)
        assert np.allclose(rho, v.a_or / gm.thermal_limit_a[None, :], rtol=1e-6)
    eng.close()


def test_rows_mode_matches_series_mode(cuda_required):
    """b200pf_run_rows_staged (float32 rows from pinned host memory) == device-resident series stepping."""
    from grid2op_b200.rollout import BatchedDoNothing
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    env = BatchedDoNothing(gm, chron, 200)
    for _ in range(3):
        env.step_device()
        o1, s1, i1, _ = env.fetch()
        o2, s2 = env.step_host()
        assert (s1 == 0).all() and np.array_equal(s1, s2)
        assert np.array_equal(o1, o2)
    env.close()


def test_precollated_host_path_matches_series_mode(cuda_required):
    from grid2op_b200.rollout import BatchedDoNothing
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    env = BatchedDoNothing(gm, chron, 96)
    assert env.precollate()
    for _ in range(4):
        env.step_device()
        o1, s1, i1, _ = env.fetch()
        o2, s2 = env.step_host()
        assert (s1 == 0).all() and np.array_equal(s1, s2) and np.array_equal(o1, o2)
    env.close()


@pytest.mark.parametrize("direct", [False, True, 2, 6])      # group flags of include/b200pf.h (True = 1)
@pytest.mark.parametrize("collated", [False, True])
def test_group_pipelined_host_path_matches_series_mode(cuda_required, collated, direct):
    """b200pf_rows_group_launch / _wait: groups stepped out of phase give, per instance and step, exactly the
    results of the device-resident series stepping."""
    from grid2op_b200.rollout import BatchedDoNothing
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    B, K = 150, 4
    ref = BatchedDoNothing(gm, chron, B)
    want = []
    for _ in range(K):
        ref.step_device()
        o, s, _, _ = ref.fetch()
        want.append((o.copy(), s.copy()))
    ref.close()
    env = BatchedDoNothing(gm, chron, B)
    if collated:
        assert env.precollate()
    groups = env.host_groups(3, direct_out=direct)
    assert groups[0][0] == 0 and groups[-1][1] == B
    # stagger the groups: group g is g steps ahead of the last one at any time
    done = [0] * len(groups)
    for g in range(len(groups)):
        env.group_launch(g)
    while min(done) < K:
        for g, (lo, hi) in enumerate(groups):
            if done[g] >= K:
                continue
            o, s = env.group_wait(g)
            wo, ws = want[done[g]]
            assert (s == 0).all() and np.array_equal(s, ws[lo:hi]) and np.array_equal(o, wo[lo:hi])
            done[g] += 1
            if done[g] < K:
                env.group_launch(g)
                if g == 0 and done[g] < K - 1:       # let group 0 run one extra step ahead
                    o, s = env.group_wait(g)
                    wo, ws = want[done[g]]
                    assert np.array_equal(o, wo[lo:hi])
                    done[g] += 1
                    env.group_launch(g)
    env.group_launch(0)
    with pytest.raises(Exception):           # a group in flight must be waited for before its next launch
        env.group_launch(0)
    env.group_wait(0)
    env.close()


@pytest.mark.parametrize("with_host_topo", [False, True])
def test_device_pointer_entry_points(cuda_required, with_host_topo):
    """b200pf_run_device (pivoting kernels: no host copy of the topology) and b200pf_run_device_topo (planned kernel) on
    torch tensors == the host-buffer entry point."""
    import torch
    from grid2op_b200.engine import PowerFlowEngine
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    z = np.load(os.path.join(GOLD, "oracle_case14_steps.npz"))
    topo, inj = np.ascontiguousarray(z["topo"], dtype=np.int8), np.ascontiguousarray(z["inj"], dtype=np.float64)
    B = len(topo)
    eng = PowerFlowEngine(gm, max_batch=B)
    want, ws, wi, _ = eng.run(topo, inj)
    dev = torch.device("cuda", 0)
    t_topo, t_inj = torch.from_numpy(topo).to(dev), torch.from_numpy(inj).to(dev)
    t_out = torch.empty((B, gm.n_out), dtype=torch.float32, device=dev)
    t_st = torch.empty(B, dtype=torch.int32, device=dev)
    t_it = torch.empty(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    eng.run_device(B, t_topo.data_ptr(), t_inj.data_ptr(), t_out.data_ptr(), t_st.data_ptr(), t_it.data_ptr(),
                   host_topo=topo if with_host_topo else None)
    eng.sync()
    assert (eng.plan_stats()["last_kernel"].startswith("planned")) == with_host_topo
    assert np.array_equal(t_st.cpu().numpy(), ws) and np.array_equal(t_it.cpu().numpy(), wi)
    _compare(gm, t_out.cpu().numpy(), want, ws == 0)
    eng.close()
