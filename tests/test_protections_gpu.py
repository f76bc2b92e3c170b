"""Device-side protections (cascading failure inside the series step) replay the reference's recorded
DoNothing rollouts of rte_case5_example (grid2op/data/rte_case5_example/_statistics: PandaPowerBackend, AC,
NB_TIMESTEP_OVERFLOW_ALLOWED=2, HARD_OVERFLOW_THRESHOLD=2): all 20 scenarios as ONE batch of 20 instances,
every step one fused kernel launch, no host logic in the loop.  Lines must trip at the same step, the game
must be over at the same step, and every observed quantity must match the recording."""
import json
import os

import numpy as np
import pytest

from conftest import grid2op_root

from grid2op_b200.gridmodel import GridModel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("policy", [0, 2])      # 0: warp kernel, cascade entirely on the device; 2: planned kernel + host re-planning
def test_batched_rollout_with_protections_matches_recording(cuda_required, policy):
    root = grid2op_root()
    stat = os.path.join(root, "data", "rte_case5_example", "_statistics") if root else None
    if stat is None or not os.path.isdir(stat):
        pytest.skip("reference rollouts not available")
    from grid2op_b200.engine import OutputView
    from grid2op_b200.rollout import BatchedDoNothing
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_rte_case5_example.npz"))
    chron = np.load(os.path.join(GOLD, "case5_chronics.npz"))["chron"]
    meta = json.load(open(os.path.join(stat, "metadata.json")))
    keys = ["p_or", "q_or", "v_or", "a_or", "p_ex", "q_ex", "v_ex", "a_ex", "prod_p", "prod_q", "prod_v", "load_v", "rho",
            "line_status", "timestep_overflow"]
    gold = {k: np.load(os.path.join(stat, f"obs_{k}.npz"))["data"] for k in keys}
    sid = np.load(os.path.join(stat, "scenario_ids.npz"))["data"].ravel().astype(int)
    rows = {s: np.flatnonzero(sid == s) for s in range(20)}
    nb_step = np.array([meta[str(s)]["nb_step"] for s in range(20)])
    B = 20
    env = BatchedDoNothing(gm, chron, B, scen=np.arange(B), t0=np.zeros(B), protections=True,
                           hard_overflow_threshold=2.0, soft_overflow_threshold=1.0, nb_timestep_overflow_allowed=2)
    env.engine.set_kernel_policy(policy)
    worst = {k: 0.0 for k in keys}
    first_done = np.full(B, -1)
    n_cmp = 0
    for k in range(int(nb_step.max())):
        if k == 0:
            env.reset_step()                              # the step env.reset() performs
        else:
            env.step_device()
        out, status, iters, rho = env.fetch()
        assert (env.engine.plan_stats()["last_kernel"].startswith("planned")) == (policy == 2)
        st = env.fetch_state()
        v = OutputView(gm, out)
        for s in range(B):
            if st["done"][s] and first_done[s] < 0:
                first_done[s] = k
            if k >= nb_step[s] - 1:                       # recorded rows end one before the game-over step
                continue
            assert status[s] == 0, (s, k, status[s])
            r = rows[s][k]
            mine = dict(p_or=v.p_or[s], q_or=v.q_or[s], v_or=v.v_or[s], a_or=v.a_or[s], p_ex=v.p_ex[s], q_ex=v.q_ex[s],
                        v_ex=v.v_ex[s], a_ex=v.a_ex[s], prod_p=v.unit_p[s], prod_q=v.unit_q[s], prod_v=v.unit_v[s],
                        load_v=v.load_v[s], rho=rho[s])
            for name, val in mine.items():
                worst[name] = max(worst[name], float(np.max(np.abs(val.astype(np.float64) - gold[name][r]))))
            assert np.array_equal(v.a_or[s] > 0, gold["line_status"][r]) or np.array_equal(
                (np.abs(v.p_or[s]) > 0) | (v.v_or[s] > 0), gold["line_status"][r]), (s, k)
            assert np.array_equal(st["timestep_overflow"][s], gold["timestep_overflow"][r]), (s, k)
            n_cmp += 1
    # game over exactly where the reference's episode ended (scenario 13 runs to the end of its chronics)
    for s in range(B):
        if nb_step[s] < chron.shape[1]:
            assert first_done[s] == nb_step[s] - 1, (s, first_done[s], nb_step[s])
    assert n_cmp > 7000
    for name in ("p_or", "q_or", "p_ex", "q_ex", "prod_p", "prod_q"):
        assert worst[name] <= 1e-4, (name, worst[name])          # sn_mva = 1 -> 1e-4 p.u.
    for name in ("v_or", "v_ex", "prod_v", "load_v"):
        assert worst[name] <= 2e-5, (name, worst[name])
    for name in ("a_or", "a_ex"):
        assert worst[name] <= 2e-3, (name, worst[name])
    assert worst["rho"] <= 1e-6
    env.close()


def test_planned_protections_on_a_large_grid_match_the_emulated_kernel(cuda_required):
    """36 substations (beyond the warp kernel): device run of the planned kernel with protections == the host build of the
    same kernel + cascade loop (tests/emu), thermal limits lowered so that lines trip and cascades happen."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from bench_configs import synth_chron
    from grid2op_b200.rollout import BatchedDoNothing
    from sparse_emu import EmuProtRollout
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_neurips_2020_track1.npz"))
    chron = synth_chron(gm, n_rows=12, seed=4)
    B = 24
    scen = np.zeros(B, dtype=np.int32)
    t0 = (np.arange(B) % chron.shape[1]).astype(np.int32)
    # limits: a base solve's flows scaled so that a few lines sit above 100 % and one or two above 200 %
    probe = BatchedDoNothing(gm, chron, B, scen=scen, t0=t0)
    probe.step_device()
    out, status, _, _ = probe.fetch()
    probe.close()
    assert (status == 0).all()
    from grid2op_b200.engine import OutputView
    a0 = OutputView(gm, out).a_or.max(axis=0)
    th = np.maximum(a0 * 1.3, 1.0).astype(np.float32)
    order = np.argsort(-a0)
    th[order[0]] = a0[order[0]] * 0.45          # hard overflow at once
    th[order[3]] = a0[order[3]] * 0.9           # soft overflow -> trips after the allowed steps
    th[order[5]] = a0[order[5]] * 0.95
    env = BatchedDoNothing(gm, chron, B, scen=scen, t0=t0, protections=True, thermal_limit_a=th,
                           hard_overflow_threshold=2.0, soft_overflow_threshold=1.0, nb_timestep_overflow_allowed=2)
    ref = EmuProtRollout(gm, chron, scen, t0, th, hard=2.0, soft=1.0, max_allowed=2)
    n_trip = 0
    for k in range(10):
        if k == 0:
            env.reset_step()
        else:
            env.step_device()
        ref.step(from_reset=(k == 0))
        assert env.engine.plan_stats()["last_kernel"].startswith("planned")
        out, status, iters, rho = env.fetch()
        st = env.fetch_state()
        assert np.array_equal(st["done"], ref.done), k
        assert np.array_equal(status, ref.status), k
        live = ref.done == 0
        assert np.array_equal(st["disc_lines"][live], ref.disc[live]), k
        assert np.array_equal(st["timestep_overflow"][live], ref.ts_over[live]), k
        assert np.array_equal(st["protection_counter"][live], ref.pcount[live]), k
        ok = live & (status == 0)
        assert np.allclose(rho[ok], ref.rho[ok], rtol=2e-5, atol=1e-6), k
        assert np.allclose(out[ok], ref.out[ok], rtol=2e-5, atol=2e-4), k
        n_trip += int((ref.disc[live] >= 0).sum())
    assert n_trip > 0
    env.close()


def test_lines_tripped_on_the_device_stay_out_when_the_planned_kernel_takes_over(cuda_required):
    """Device-side cascade of the warp kernel writes the outages into the device copy of the topology only; switching the
    protections off (-> planned kernel, host-side plans) must not bring the lines back."""
    from grid2op_b200.engine import OutputView
    from grid2op_b200.rollout import BatchedDoNothing
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    B = 64
    probe = BatchedDoNothing(gm, chron, B)
    probe.step_device()
    out, status, _, _ = probe.fetch()
    probe.close()
    a_all = OutputView(gm, out).a_or
    a0 = a_all.max(axis=0)
    th = np.maximum(a0 * 1.5, 1.0).astype(np.float32)
    victim = int(np.argsort(-a0)[2])
    th[victim] = a_all[:, victim].min() * 0.4                                 # far above its limit in EVERY instance: trips at once
    env = BatchedDoNothing(gm, chron, B, protections=True, thermal_limit_a=th)
    env.reset_step()
    env.step_device()
    out, status, _, _ = env.fetch()
    assert env.engine.plan_stats()["last_kernel"] == "warp_pivoting"
    live = status == 0
    assert live.sum() >= B // 2
    off = OutputView(gm, out).a_or == 0.0
    assert off[live].any(axis=1).all()                                         # every live instance lost at least that line
    env.engine.series_protections(False)
    env.step_device()
    out2, status2, _, _ = env.fetch()
    assert env.engine.plan_stats()["last_kernel"].startswith("planned")
    both = live & (status2 == 0)
    off2 = OutputView(gm, out2).a_or == 0.0
    assert np.array_equal(off2[both], off[both])
    env.close()
