"""``BatchedEnv`` (grid2op_b200/batched_env.py: batched env.step WITH actions, cooldowns, maintenance) against unmodified
grid2op environments stepped one by one with the same scripted agent, on l2rpn_neurips_2020_track1 (36 substations; its
chronics carry a maintenance table): topology vectors, line status and both cooldown vectors must be EQUAL after every step,
rho within tolerance, illegal actions counted the same.

* CPU (``not gpu``): host logic only — the reference environments run B200Backend with the oracle adapter as engine, the
  batched driver runs on the oracle's C restatement (tests/oracle_engine.py: test infrastructure), protections off;
* GPU (``-m gpu``): the real CUDA engine on both sides, protections ON (Backend.next_grid_state on the device vs the
  reference's host loop), a longer run that crosses the end of the maintenance.
"""
import os
import warnings

import numpy as np
import pytest

from conftest import env_grid, have_cuda

ENV = "l2rpn_neurips_2020_track1"


def _scripted(k, rng, obs_topo, line_status, line_cd, sub_cd, sub_pos, sub_size, n_line, is_line_pos):
    """-> dict(sub=-1|s, bus=[...], line=-1|l, status=0|+-1): what the agent does at step k (may be illegal on purpose).
    Substation actions put two of the connected line ends on busbar 2 (a pass-through node: nothing gets isolated) or
    everything back on busbar 1."""
    act = dict(sub=-1, bus=None, line=-1, status=0)
    if k % 3 == 0 or k % 10 == 4:
        s = int(rng.integers(0, len(sub_size)))
        if k % 10 == 4 and (sub_cd > 0).any():
            s = int(np.flatnonzero(sub_cd > 0)[0])          # on purpose: a substation that is still in cooldown -> illegal
        pos = sub_pos[s, :sub_size[s]]
        live = obs_topo[pos] > 0
        bus = np.where(live, 1, 0)
        ends = np.flatnonzero(live & is_line_pos[pos])
        if len(ends) >= 4 and rng.random() < 0.7:
            bus[rng.choice(ends, 2, replace=False)] = 2
        if (bus > 0).any():
            act.update(sub=s, bus=bus)
    elif k % 5 == 1:
        l = int(rng.integers(0, n_line))
        act.update(line=l, status=-1 if line_status[l] else +1)
    return act


def _run(n_steps, start_row, backend_factory, engine_factory, protections, seeds=(0, 1, 2), scen_name="Scenario_august_dummy"):
    import grid2op
    from grid2op.Parameters import Parameters
    from grid2op_b200.batched_env import BatchedEnv
    from grid2op_b200.chronics import load_line_events, load_scenarios
    from grid2op_b200.gridmodel import GridModel
    grid = env_grid(ENV)
    gm = GridModel(grid)
    cdir = os.path.join(os.path.dirname(grid), "chronics")
    folder = os.path.join(cdir, scen_name)
    chron = load_scenarios(cdir, gm, scenarios=[folder])
    maint = load_line_events(folder, gm, "maintenance")[None]
    B = len(seeds)
    p = Parameters()
    p.NO_OVERFLOW_DISCONNECTION = not protections
    p.NB_TIMESTEP_COOLDOWN_LINE = 3
    p.NB_TIMESTEP_COOLDOWN_SUB = 2
    envs = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(B):
            e = grid2op.make(ENV, test=True, backend=backend_factory(), param=p, opponent_init_budget=0., opponent_budget_per_ts=0.,
                             _add_to_name=f"benv{i}_{int(protections)}")
            sid = [j for j, f in enumerate(sorted(os.listdir(cdir))) if f == scen_name][0]
            e.set_id(sid)
            e.reset()
            e.fast_forward_chronics(start_row)
            envs.append(e)
    obs = [e.get_obs() for e in envs]
    th = np.asarray(envs[0].get_thermal_limit(), dtype=np.float32)
    benv = BatchedEnv(gm, chron, B, maintenance=maint, scen=np.zeros(B, dtype=np.int32), t0=np.full(B, start_row + 1, dtype=np.int32),
                      thermal_limit_a=th, nb_timestep_cooldown_line=3, nb_timestep_cooldown_sub=2,
                      nb_timestep_reconnection=p.NB_TIMESTEP_RECONNECTION, protections=protections,
                      hard_overflow_threshold=p.HARD_OVERFLOW_THRESHOLD, soft_overflow_threshold=p.SOFT_OVERFLOW_THRESHOLD,
                      nb_timestep_overflow_allowed=p.NB_TIMESTEP_OVERFLOW_ALLOWED, engine=engine_factory(gm))
    # bring the batched state to the reference's state after fast_forward (same topology: everything on busbar 1)
    for i in range(B):
        assert np.array_equal(obs[i].topo_vect, benv.topo[i, :gm.dim_topo])
        benv.line_cooldown[i] = obs[i].time_before_cooldown_line
        benv.sub_cooldown[i] = obs[i].time_before_cooldown_sub
    rngs = [np.random.default_rng(100 + s) for s in seeds]
    is_line_pos = np.zeros(gm.dim_topo, dtype=bool)
    is_line_pos[gm.line_or_pos] = True; is_line_pos[gm.line_ex_pos] = True
    n_illegal_ref = 0
    saw = dict(maint=False, illegal=False, sub=False, line=False)
    alive = np.ones(B, dtype=bool)
    for k in range(n_steps):
        sub_id = np.full(B, -1, dtype=np.int64); sub_bus = np.zeros((B, benv.max_sub_size), dtype=np.int8)
        line_id = np.full(B, -1, dtype=np.int64); line_st = np.zeros(B, dtype=np.int64)
        infos = []
        for i, e in enumerate(envs):
            if not alive[i]:
                infos.append(None)
                continue
            o = obs[i]
            a = _scripted(k, rngs[i], o.topo_vect, o.line_status, o.time_before_cooldown_line, o.time_before_cooldown_sub,
                          benv.sub_pos, benv.sub_size, gm.n_line, is_line_pos)
            spec = {}
            if a["sub"] >= 0:
                spec["set_bus"] = {"substations_id": [(a["sub"], a["bus"].astype(int).tolist())]}
                sub_id[i] = a["sub"]; sub_bus[i, :len(a["bus"])] = a["bus"]
                saw["sub"] = True
            if a["line"] >= 0:
                spec["set_line_status"] = [(a["line"], int(a["status"]))]
                line_id[i] = a["line"]; line_st[i] = a["status"]
                saw["line"] = True
            o2, r, d, info = e.step(e.action_space(spec))
            obs[i] = o2
            infos.append((d, info))
            if info["is_illegal"]:
                n_illegal_ref += 1; saw["illegal"] = True
        rho, done, binfo = benv.step(sub_id, sub_bus, line_id, line_st)
        for i in range(B):
            if not alive[i]:
                continue
            d, info = infos[i]
            assert bool(done[i]) == bool(d), (k, i, info["exception"], binfo["status"][i])
            if d:
                alive[i] = False
                continue
            o = obs[i]
            assert np.array_equal(o.topo_vect, benv.topo[i, :gm.dim_topo]), (k, i, np.flatnonzero(o.topo_vect != benv.topo[i, :gm.dim_topo]))
            assert np.array_equal(o.line_status, benv.line_status()[i]), (k, i)
            assert np.array_equal(o.time_before_cooldown_line, benv.line_cooldown[i]), (k, i, o.time_before_cooldown_line, benv.line_cooldown[i])
            assert np.array_equal(o.time_before_cooldown_sub, benv.sub_cooldown[i]), (k, i)
            assert np.allclose(o.rho, rho[i], rtol=2e-4, atol=2e-5), (k, i, float(np.max(np.abs(o.rho - rho[i]))))
            if (o.time_next_maintenance == 0).any():
                saw["maint"] = True
    assert benv.n_illegal == n_illegal_ref
    saw["steps_alive"] = int(benv.n_steps)
    saw["n_alive"] = int(alive.sum())
    for e in envs:
        e.close()
    benv.close()
    return saw


def test_batched_env_host_logic_vs_reference_env():
    if env_grid(ENV) is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    from oracle_engine import COracleSeriesEngine, OracleEngine

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    saw = _run(24, 96, HostLogicBackend, COracleSeriesEngine, protections=False, seeds=(0, 1))
    assert saw["maint"] and saw["sub"] and saw["line"]


@pytest.mark.gpu
@pytest.mark.parametrize("protections", [False, True])
def test_batched_env_vs_reference_env_gpu(cuda_required, protections):
    if env_grid(ENV) is None:
        pytest.skip("reference data not available")
    from grid2op_b200.backend import B200Backend
    saw = _run(130, 96, B200Backend, lambda gm: None, protections=protections, seeds=(0, 1, 2, 3))
    assert saw["maint"] and saw["sub"] and saw["line"] and saw["illegal"]


@pytest.mark.gpu
def test_reset_of_single_instances(cuda_required):
    """game over -> BatchedEnv.reset_instances: flags / counters cleared on the device too, the instances step again"""
    from grid2op_b200.batched_env import BatchedEnv
    from grid2op_b200.engine import OutputView
    from grid2op_b200.gridmodel import GridModel
    from grid2op_b200.rollout import BatchedDoNothing
    gold = os.path.join(os.path.dirname(__file__), "golden")
    gm = GridModel.from_npz(os.path.join(gold, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(gold, "case14_sandbox_chronics.npz"))["chron"]
    B = 64
    probe = BatchedDoNothing(gm, chron, B)
    probe.step_device()
    out, _, _, _ = probe.fetch()
    probe.close()
    a_all = OutputView(gm, out).a_or
    th = np.maximum(a_all.max(axis=0) * 1.5, 1.0).astype(np.float32)
    env = BatchedEnv(gm, chron, B, protections=True, thermal_limit_a=th)
    env.reset_step()
    # make half of the instances game over: disconnect lines until the grid falls apart (scripted through the action API)
    victims = np.arange(0, B, 2)
    for l in range(gm.n_line):
        lid = np.full(B, -1); lst = np.zeros(B, dtype=np.int64)
        lid[victims] = l; lst[victims] = -1
        rho, done, info = env.step(line_id=lid, line_status=lst)
        if done[victims].all():
            break
    assert done[victims].all() and not done[1::2].any()
    st = env.engine.series_fetch_state()
    assert (st["done"][victims] == 1).all()
    env.reset_instances(victims)
    st = env.engine.series_fetch_state()
    assert (st["done"] == 0).all() and (st["protection_counter"][victims] == 0).all() and (st["disc_lines"][victims] == -1).all()
    rho, done, info = env.step()
    assert not done.any() and (info["status"] == 0).all()
    assert np.isfinite(rho).all()
    env.close()


def test_reset_of_single_instances_host_logic():
    """reset_instances on the oracle stand-in engine: the reset instance replays exactly what a fresh driver produces, the others go on"""
    from grid2op_b200.batched_env import BatchedEnv
    from grid2op_b200.gridmodel import GridModel
    from oracle_engine import COracleSeriesEngine
    gold = os.path.join(os.path.dirname(__file__), "golden")
    gm = GridModel.from_npz(os.path.join(gold, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(gold, "case14_sandbox_chronics.npz"))["chron"]
    B = 4
    t0 = np.array([3, 3, 7, 7], dtype=np.int32)
    mk = lambda: BatchedEnv(gm, chron, B, scen=np.zeros(B, dtype=np.int32), t0=t0, nb_timestep_cooldown_line=3,        # noqa: E731
                            engine=COracleSeriesEngine(gm))
    env, fresh = mk(), mk()
    first, _, _ = fresh.step()
    lid = np.array([2, -1, 5, -1]); lst = np.array([-1, 0, -1, 0])
    rho1, _, _ = env.step(line_id=lid, line_status=lst)
    assert rho1[0, 2] == 0 and env.line_cooldown[0, 2] == 3
    rho2, _, _ = env.step()
    env.reset_instances([0, 2])
    assert (env.line_cooldown[[0, 2]] == 0).all() and (env.topo[[0, 2]] == env.topo0[[0, 2]]).all()
    assert env.row[0] == 3 and env.row[2] == 7 and env.row[1] == 5
    rho3, done, _ = env.step()
    assert not done.any()
    assert np.array_equal(rho3[[0, 2]], first[[0, 2]])                 # a fresh episode from the same rows
    third, _, _ = (fresh.step(), fresh.step())[1]
    assert np.array_equal(rho3[[1, 3]], third[[1, 3]])                 # the others: third step of an undisturbed run


def test_bus_cap_follows_the_topology():
    """BatchedEnv.nb_cap = active buses of the fullest instance (exact up to 17, then rounded up to a multiple of 8)"""
    from grid2op_b200.batched_env import BatchedEnv
    from grid2op_b200.gridmodel import GridModel
    from grid2op_b200.rollout import BatchedDoNothing
    from oracle_engine import COracleSeriesEngine
    gold = os.path.join(os.path.dirname(__file__), "golden")
    for name, chr_ in (("l2rpn_case14_sandbox", "case14_sandbox_chronics.npz"), ("l2rpn_neurips_2020_track1", "neurips_2020_track1_chronics.npz")):
        gm = GridModel.from_npz(os.path.join(gold, f"gridmodel_{name}.npz"))
        chron = np.load(os.path.join(gold, chr_))["chron"]
        env = BatchedEnv(gm, chron, 3, engine=COracleSeriesEngine(gm))
        plain = BatchedDoNothing(gm, chron, 3, engine=COracleSeriesEngine(gm)).nb_cap
        rnd = lambda n: n if n <= 17 else min((n + 7) // 8 * 8, gm.n_slot)          # noqa: E731
        assert env.nb_cap == rnd(plain)
        # split substation 1 of instance 2 (two elements to busbar 2): one more active bus
        bus = np.zeros((3, env.max_sub_size), dtype=np.int8)
        bus[2, :env.sub_size[1]] = 1; bus[2, :2] = 2
        env.step(sub_id=np.array([-1, -1, 1]), sub_bus=bus)
        assert env.nb_cap == rnd(plain + 1)
        # brute force on the topology records
        def count(tv):
            return len(set(env._slots(tv).tolist()))
        assert max(count(env.topo[i]) for i in range(3)) == plain + 1
        env.reset_instances([2])
        assert env.nb_cap == rnd(plain)
        off = BatchedEnv(gm, chron, 3, engine=COracleSeriesEngine(gm), tight_cap=False)
        assert off.nb_cap == 0


@pytest.mark.gpu
def test_pivoting_kernels_with_tight_cap_equal_planned(cuda_required):
    """random-walk topologies (BASELINE configs[2]) through the kernels that discover the topology on the device, workspace sized by
    BatchedEnv's bus cap, against the planned kernel (plans built on the host) and against the same kernels sized for every bus
    slot: same statuses (no ST_LARGE ever), same flows"""
    from grid2op_b200.batched_env import BatchedEnv, random_substation_actions
    from grid2op_b200.gridmodel import GridModel
    gold = os.path.join(os.path.dirname(__file__), "golden")
    gm = GridModel.from_npz(os.path.join(gold, "gridmodel_l2rpn_neurips_2020_track1.npz"))
    chron = np.load(os.path.join(gold, "neurips_2020_track1_chronics.npz"))["chron"]
    B = 192
    envs = [BatchedEnv(gm, chron, B), BatchedEnv(gm, chron, B, tight_cap=False), BatchedEnv(gm, chron, B)]
    for e, pol in zip(envs, (1, 1, 2)):
        e.engine.set_kernel_policy(pol)
        e.reset_step()
    rng = np.random.default_rng(5)
    caps = []
    n_conv = 0
    for k in range(10):
        sub, bus = random_substation_actions(envs[0], rng)
        res = [e.step(sub, bus) for e in envs]
        caps.append(envs[0].nb_cap)
        st = [r[2]["status"] for r in res]
        assert not (st[0] == 4).any() and not (st[1] == 4).any()                      # ST_LARGE
        assert np.array_equal(st[0], st[1])
        assert int((st[0] != st[2]).sum()) <= 1, (k, np.flatnonzero(st[0] != st[2]))   # (pivoting vs planned + safety net: borderline states)
        ok = (st[0] == 0) & (st[2] == 0) & ~res[0][1] & ~res[2][1]
        n_conv += int(ok.sum())
        assert np.allclose(res[0][0][ok], res[1][0][ok], rtol=1e-5, atol=1e-6)
        assert np.allclose(res[0][0][ok], res[2][0][ok], rtol=2e-3, atol=2e-4), float(np.max(np.abs(res[0][0][ok] - res[2][0][ok])))
        done = res[0][1] | res[2][1]
        if done.any():
            who = np.flatnonzero(done)
            for e in envs:
                e.reset_instances(who, rows=e.row[who])
    assert envs[0].engine.plan_stats()["last_kernel"] != "planned_sparse" and envs[2].engine.plan_stats()["last_kernel"].startswith("planned")
    assert 36 < max(caps) < gm.n_slot          # above the plain topology's bus count, below "every slot"
    assert n_conv > 5 * B
    for e in envs:
        e.close()


def test_busbar_left_with_an_out_of_service_shunt_only_host_logic():
    """l2rpn_neurips_2020_track1: the grid file has its six shunts out of service, the reference's environment still counts busbar 1 of
    their substations as active (``active_bus``, _backendAction.py:1519-1531 -> bus.in_service, pPB:920-922).  Moving EVERY element of
    such a substation to busbar 2 leaves an in-service bus without a voltage: the reference ends the episode ("Isolated bus",
    pPB:1241-1244); the same move on a substation without a shunt is an ordinary (pass-through) topology.  BatchedEnv and unmodified
    environments (B200Backend host logic) must agree on both."""
    if env_grid(ENV) is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    import grid2op
    from grid2op.Parameters import Parameters
    from oracle_engine import COracleSeriesEngine, OracleEngine
    from grid2op_b200.batched_env import BatchedEnv
    from grid2op_b200.chronics import load_scenarios
    from grid2op_b200.gridmodel import GridModel

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    grid = env_grid(ENV)
    gm = GridModel(grid)
    cdir = os.path.join(os.path.dirname(grid), "chronics")
    folder = os.path.join(cdir, sorted(os.listdir(cdir))[0])
    chron = load_scenarios(cdir, gm, scenarios=[folder])
    shunt_subs = set(int(x) for x in gm.shunt_sub)
    p = Parameters()
    p.NO_OVERFLOW_DISCONNECTION = True
    envs = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(2):
            e = grid2op.make(ENV, test=True, backend=HostLogicBackend(), param=p, opponent_init_budget=0., opponent_budget_per_ts=0.,
                             _add_to_name=f"benv_iso{i}")
            e.set_id(0)
            e.reset()
            envs.append(e)
    th = np.asarray(envs[0].get_thermal_limit(), dtype=np.float32)
    benv = BatchedEnv(gm, chron, 2, scen=np.zeros(2, dtype=np.int32), t0=np.full(2, 1, dtype=np.int32), thermal_limit_a=th,
                      protections=False, engine=COracleSeriesEngine(gm))
    s_shunt = 12
    assert s_shunt in shunt_subs and benv.sub_size[s_shunt] == 4
    s_plain = [s for s in range(gm.n_sub) if s not in shunt_subs and 3 <= benv.sub_size[s] <= 5 and not np.isin(s, gm.gen_sub)]
    assert s_plain
    subs = [s_shunt, s_plain[0]]
    sub_id = np.asarray(subs, dtype=np.int64)
    sub_bus = np.zeros((2, benv.max_sub_size), dtype=np.int8)
    ref_done = []
    for i, s in enumerate(subs):
        n = int(benv.sub_size[s])
        sub_bus[i, :n] = 2
        o, r, d, info = envs[i].step(envs[i].action_space({"set_bus": {"substations_id": [(s, [2] * n)]}}))
        ref_done.append(bool(d))
    rho, done, info = benv.step(sub_id, sub_bus, np.full(2, -1), np.zeros(2, dtype=np.int64))
    assert ref_done == [True, False]
    assert done.tolist() == ref_done, (done, info["status"])
    for e in envs:
        e.close()
    benv.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_substation_actions_host_logic(seed):
    """BASELINE configs[2] as a parity test: every instance draws one substation and a random busbar for each of its elements, every
    step (``random_substation_actions``), side by side with unmodified environments fed the same ``set_bus`` actions — game overs
    (isolated elements, islanded grids, the isolated-busbar rule) must coincide step by step, topology vectors / cooldowns equal,
    rho within tolerance.  CPU: host logic on both sides (oracle adapters)."""
    if env_grid(ENV) is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    import grid2op
    from grid2op.Parameters import Parameters
    from oracle_engine import COracleSeriesEngine, OracleEngine
    from grid2op_b200.batched_env import BatchedEnv, random_substation_actions
    from grid2op_b200.chronics import load_scenarios
    from grid2op_b200.gridmodel import GridModel

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    grid = env_grid(ENV)
    gm = GridModel(grid)
    cdir = os.path.join(os.path.dirname(grid), "chronics")
    folder = os.path.join(cdir, sorted(os.listdir(cdir))[0])
    chron = load_scenarios(cdir, gm, scenarios=[folder])
    B = 6
    p = Parameters()
    p.NO_OVERFLOW_DISCONNECTION = True
    p.NB_TIMESTEP_COOLDOWN_SUB = 1
    envs = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(B):
            e = grid2op.make(ENV, test=True, backend=HostLogicBackend(), param=p, opponent_init_budget=0., opponent_budget_per_ts=0.,
                             _add_to_name=f"benv_rnd{i}")
            e.set_id(0)
            e.reset()
            envs.append(e)
    th = np.asarray(envs[0].get_thermal_limit(), dtype=np.float32)
    benv = BatchedEnv(gm, chron, B, scen=np.zeros(B, dtype=np.int32), t0=np.full(B, 1, dtype=np.int32), thermal_limit_a=th,
                      nb_timestep_cooldown_sub=1, protections=False, engine=COracleSeriesEngine(gm))
    rng = np.random.default_rng(50 + seed)
    alive = np.ones(B, dtype=bool)
    n_over = n_checked = 0
    for k in range(10):
        sub, bus = random_substation_actions(benv, rng)
        ref = []
        for i, e in enumerate(envs):
            if not alive[i]:
                ref.append(None)
                continue
            n = int(benv.sub_size[sub[i]])
            o, r, d, info = e.step(e.action_space({"set_bus": {"substations_id": [(int(sub[i]), bus[i, :n].astype(int).tolist())]}}))
            ref.append((o, d, info))
        rho, done, binfo = benv.step(sub, bus)
        for i in range(B):
            if not alive[i]:
                continue
            o, d, info = ref[i]
            assert bool(done[i]) == bool(d), (k, i, int(sub[i]), bus[i], info["exception"], binfo["status"][i])
            if d:
                alive[i] = False; n_over += 1
                continue
            n_checked += 1
            assert np.array_equal(o.topo_vect, benv.topo[i, :gm.dim_topo]), (k, i)
            assert np.array_equal(o.time_before_cooldown_sub, benv.sub_cooldown[i]), (k, i)
            assert np.allclose(o.rho, rho[i], rtol=2e-4, atol=2e-5), (k, i, float(np.max(np.abs(o.rho - rho[i]))))
        if not alive.any():
            break
    assert n_checked >= 6 and n_over >= 1, (n_checked, n_over)
    for e in envs:
        e.close()
    benv.close()


@pytest.mark.parametrize("seed,protections", [(0, False), (1, False), (3, False), (0, True), (2, True), (5, True)])
def test_random_mixed_actions_host_logic(seed, protections):
    if env_grid(ENV) is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    from oracle_engine import COracleSeriesEngine, EmuProtSeriesEngine, OracleEngine

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    run_random_mixed_actions(seed, protections, HostLogicBackend, EmuProtSeriesEngine if protections else COracleSeriesEngine)


def run_random_mixed_actions(seed, protections, backend_class, engine_factory):
    """(``protections``: with the overflow protections of Backend.next_grid_state, backend.py:1433-1521 — the batched side runs the HOST
    BUILD of the planned kernel with its cascade loop, tests/oracle_engine.py EmuProtSeriesEngine; tripped lines, their reconnection
    cooldown and the cascades must coincide with the reference's host loop.)
    Random mix of substation re-assignments (a busbar for EVERY element, ends of disconnected lines included), line switching and
    do-nothing, with line / substation cooldowns: a busbar > 0 for the end of a disconnected line is a reconnection of that line
    (impact rules of grid2op/Action/baseAction.py:1836-1860: the line counts, its ends do not count for the substation; illegal while
    the line is in cooldown or when more than MAX_LINE_STATUS_CHANGED lines are touched; the other end returns to its last busbar).
    Topology vectors, both cooldown vectors, game overs and rho equal to unmodified environments after every step."""
    import grid2op_b200.backend as bk           # noqa: F401  (locates / bootstraps the grid2op install first)
    import grid2op
    from grid2op.Parameters import Parameters
    from grid2op_b200.batched_env import BatchedEnv, random_substation_actions
    from grid2op_b200.chronics import load_scenarios
    from grid2op_b200.engine import OutputView
    from grid2op_b200.gridmodel import GridModel

    grid = env_grid(ENV)
    gm = GridModel(grid)
    cdir = os.path.join(os.path.dirname(grid), "chronics")
    folder = os.path.join(cdir, sorted(os.listdir(cdir))[0])
    chron = load_scenarios(cdir, gm, scenarios=[folder])
    B = 5
    p = Parameters()
    p.NO_OVERFLOW_DISCONNECTION = not protections
    p.NB_TIMESTEP_COOLDOWN_SUB = 1
    p.NB_TIMESTEP_COOLDOWN_LINE = 2
    envs = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(B):
            e = grid2op.make(ENV, test=True, backend=backend_class(), param=p, opponent_init_budget=0., opponent_budget_per_ts=0.,
                             _add_to_name=f"benv_mix{i}_{int(protections)}")
            e.set_id(0)
            e.reset()
            envs.append(e)
    th = np.asarray(envs[0].get_thermal_limit(), dtype=np.float32)
    benv = BatchedEnv(gm, chron, B, scen=np.zeros(B, dtype=np.int32), t0=np.full(B, 1, dtype=np.int32), thermal_limit_a=th,
                      nb_timestep_cooldown_sub=1, nb_timestep_cooldown_line=2, nb_timestep_reconnection=p.NB_TIMESTEP_RECONNECTION,
                      protections=protections, hard_overflow_threshold=p.HARD_OVERFLOW_THRESHOLD,
                      soft_overflow_threshold=p.SOFT_OVERFLOW_THRESHOLD, nb_timestep_overflow_allowed=p.NB_TIMESTEP_OVERFLOW_ALLOWED,
                      engine=engine_factory(gm))
    rng = np.random.default_rng(seed)
    alive = np.ones(B, dtype=bool)
    n_illegal_ref = n_checked = n_open = n_trip = 0
    for k in range(14):
        sub, bus = random_substation_actions(benv, rng)
        line = np.full(B, -1, dtype=np.int64); lst = np.zeros(B, dtype=np.int64)
        kind = rng.random(B)
        ref = []
        for i, e in enumerate(envs):
            if not alive[i]:
                ref.append(None)
                continue
            spec = {}
            if kind[i] < 0.5:
                n = int(benv.sub_size[sub[i]])
                spec["set_bus"] = {"substations_id": [(int(sub[i]), bus[i, :n].astype(int).tolist())]}
            else:
                sub[i] = -1
                if kind[i] < 0.8:
                    l = int(rng.integers(0, gm.n_line))
                    line[i] = l; lst[i] = -1 if benv.line_status()[i, l] else 1
                    spec["set_line_status"] = [(l, int(lst[i]))]
            o, r, d, info = e.step(e.action_space(spec))
            n_illegal_ref += int(bool(info["is_illegal"]))
            ref.append((o, d, info, spec))
        rho, done, binfo = benv.step(sub, bus, line, lst)
        if binfo["disc_lines"] is not None:
            n_trip += int((binfo["disc_lines"] >= 0).sum())
        ts_over = benv.fetch_state()["timestep_overflow"] if protections else None
        th = None
        for i in range(B):
            if not alive[i]:
                continue
            o, d, info, spec = ref[i]
            assert bool(done[i]) == bool(d), (k, i, spec, info["exception"], binfo["status"][i])
            if d:
                alive[i] = False
                continue
            n_checked += 1
            assert np.array_equal(o.topo_vect, benv.topo[i, :gm.dim_topo]), (k, i, spec, info["is_illegal"])
            assert np.array_equal(o.time_before_cooldown_line, benv.line_cooldown[i]), (k, i, spec)
            assert np.array_equal(o.time_before_cooldown_sub, benv.sub_cooldown[i]), (k, i, spec)
            assert np.allclose(o.rho, rho[i], rtol=2e-4, atol=2e-5), (k, i)
            if protections:
                assert np.array_equal(o.timestep_overflow, ts_over[i]), (k, i)
            if th is None:
                rec = benv.fetch()[0]
                th = benv.line_angles(rec)
                view = OutputView(gm, rec)
            # the full result record of the batched step against the observation of the unmodified environment
            for a_ref, a_rec, tol in ((o.p_or, view.p_or[i], 2e-3), (o.q_or, view.q_or[i], 2e-3), (o.p_ex, view.p_ex[i], 2e-3),
                                      (o.v_or, view.v_or[i], 2e-3), (o.v_ex, view.v_ex[i], 2e-3), (o.gen_p, view.unit_p[i, gm.n_hidden:], 2e-3),
                                      (o.gen_q, view.unit_q[i, gm.n_hidden:], 2e-3), (o.gen_v, view.unit_v[i, gm.n_hidden:], 2e-3),
                                      (o.load_v, view.load_v[i], 2e-3)):
                assert np.allclose(a_ref, a_rec, rtol=2e-4, atol=tol), (k, i, float(np.max(np.abs(a_ref - a_rec))))
            assert np.allclose(o.a_or, view.a_or[i], rtol=2e-4, atol=0.05), (k, i)
            # angles as an observation shows them, open lines included (the angle of the bus the end was last attached to)
            assert np.allclose(o.theta_or, th[0][i], atol=2e-3) and np.allclose(o.theta_ex, th[1][i], atol=2e-3), (k, i, spec)
            n_open += int((~o.line_status).sum())
        if not alive.any():
            break
    assert n_open > 0 and (n_trip > 0 or not protections), (n_open, n_trip)
    assert benv.n_illegal == n_illegal_ref and n_checked >= 8
    for e in envs:
        e.close()
    benv.close()


def test_hazards_and_maintenance_tables_host_logic(tmp_path):
    """Unplanned outages (``hazards``) next to a planned one (``maintenance``): a copy of l2rpn_case14_sandbox gets synthetic tables
    written into its chronics folder (the bundled data has no hazards), an unmodified environment reads them through
    ``GridStateFromFile`` (grid2op/Chronics/gridStateFromFile.py:780-797), BatchedEnv through ``chronics.load_line_events``: forced
    outages, the cooldowns they impose (baseEnv.py:2566-2597) and the illegal reconnection attempts during them must agree."""
    import bz2
    import shutil
    if env_grid("l2rpn_case14_sandbox") is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    import grid2op
    from grid2op.Parameters import Parameters
    from oracle_engine import COracleSeriesEngine, OracleEngine
    from grid2op_b200.batched_env import BatchedEnv
    from grid2op_b200.chronics import load_line_events, load_scenarios
    from grid2op_b200.gridmodel import GridModel

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    dst = os.path.join(str(tmp_path), "l2rpn_case14_sandbox")
    shutil.copytree(os.path.dirname(env_grid("l2rpn_case14_sandbox")), dst)
    cdir = os.path.join(dst, "chronics")
    scens = sorted(os.listdir(cdir))
    folder = os.path.join(cdir, scens[0])
    for other in scens[1:]:
        shutil.rmtree(os.path.join(cdir, other))
    gm = GridModel(os.path.join(dst, "grid.json"))
    with bz2.open(os.path.join(folder, "load_p.csv.bz2"), "rt") as f:
        n_rows = sum(1 for _ in f) - 1
    haz = np.zeros((n_rows, gm.n_line), dtype=int)
    haz[5:9, 3] = 1; haz[12:14, 7] = 1; haz[13:20, 3] = 1
    mnt = np.zeros((n_rows, gm.n_line), dtype=int)
    mnt[25:30, 10] = 1
    for nm, arr in (("hazards", haz), ("maintenance", mnt)):
        with bz2.open(os.path.join(folder, nm + ".csv.bz2"), "wt") as f:
            f.write(";".join(gm.name_line) + "\n")
            for r in arr:
                f.write(";".join(str(int(x)) for x in r) + "\n")
    p = Parameters()
    p.NO_OVERFLOW_DISCONNECTION = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e = grid2op.make(dst, backend=HostLogicBackend(), param=p, _add_to_name="haz")
    e.set_id(0)
    e.reset()
    chron = load_scenarios(cdir, gm, scenarios=[folder])
    H = load_line_events(folder, gm, "hazards")[None]
    M = load_line_events(folder, gm, "maintenance")[None]
    th = np.asarray(e.get_thermal_limit(), dtype=np.float32)
    benv = BatchedEnv(gm, chron, 1, maintenance=M, hazards=H, scen=np.zeros(1, dtype=np.int32), t0=np.full(1, 1, dtype=np.int32),
                      thermal_limit_a=th, nb_timestep_reconnection=p.NB_TIMESTEP_RECONNECTION, protections=False,
                      engine=COracleSeriesEngine(gm))
    attempts = {7: 3, 10: 3, 15: 7, 22: 3, 27: 10, 31: 10}          # reconnection attempts: during the outages (illegal) and after them
    n_illegal_ref = 0
    saw_out = set()
    for k in range(40):
        lid = np.full(1, -1); lst = np.zeros(1, dtype=np.int64); spec = {}
        if k in attempts:
            lid[0] = attempts[k]; lst[0] = 1; spec = {"set_line_status": [(attempts[k], 1)]}
        o, r, d, info = e.step(e.action_space(spec))
        n_illegal_ref += int(bool(info["is_illegal"]))
        rho, done, binfo = benv.step(line_id=lid, line_status=lst)
        assert bool(done[0]) == bool(d) and not d, (k, info["exception"])
        assert np.array_equal(o.topo_vect, benv.topo[0, :gm.dim_topo]), (k, np.flatnonzero(o.topo_vect != benv.topo[0, :gm.dim_topo]))
        assert np.array_equal(o.time_before_cooldown_line, benv.line_cooldown[0]), (k, o.time_before_cooldown_line, benv.line_cooldown[0])
        assert np.allclose(o.rho, rho[0], rtol=2e-4, atol=2e-5), k
        saw_out |= set(np.flatnonzero(~o.line_status).tolist())
    assert benv.n_illegal == n_illegal_ref and n_illegal_ref >= 2 and {3, 7, 10} <= saw_out
    e.close()
    benv.close()
