"""``b200pf_grid_max_active_buses`` (host code of the library, include/b200pf.h: the tight ``nb_cap`` of a launch) against a brute-force
count of the distinct (substation, busbar) pairs that carry a connected element, on random topology records of the four grids."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["rte_case5_example", "l2rpn_case14_sandbox", "l2rpn_neurips_2020_track1", "l2rpn_wcci_2022_dev"])
def test_bus_count_vs_brute_force(name):
    from grid2op_b200.engine import grid_max_active_buses
    from grid2op_b200.gridmodel import GridModel
    gm = GridModel.from_npz(os.path.join(GOLD, f"gridmodel_{name}.npz"))
    rng = np.random.default_rng(3)
    B = 200
    topo = np.tile(gm.default_topo(), (B, 1))
    topo[:, :gm.dim_topo] = rng.integers(1, gm.n_busbar + 1, (B, gm.dim_topo))
    topo[rng.random((B, gm.n_topo_in)) < 0.1] = -1                  # disconnected elements (shunts / hidden units included)
    topo[0] = gm.default_topo()
    topo[1] = -1                                                     # nothing connected
    subs = np.concatenate([gm.line_or_sub, gm.line_ex_sub, gm.gen_sub, gm.load_sub, gm.storage_sub, gm.shunt_sub, gm.hidden_sub])
    pos = np.concatenate([gm.line_or_pos, gm.line_ex_pos, gm.gen_pos, gm.load_pos, gm.storage_pos,
                          gm.dim_topo + np.arange(gm.n_shunt), gm.dim_topo + gm.n_shunt + np.arange(gm.n_hidden)]).astype(np.int64)
    want = np.empty(B, dtype=np.int32)
    for i in range(B):
        b = topo[i, pos].astype(np.int64)
        want[i] = len(set((subs[b > 0] * 16 + b[b > 0]).tolist()))
    best, cnt = grid_max_active_buses(gm, topo)
    assert np.array_equal(cnt, want)
    assert best == want.max() and cnt[1] == 0 and cnt[0] == len(set(subs.tolist()))
    assert grid_max_active_buses(gm, topo[:0])[0] == 0


def test_batched_env_keeps_the_cap_incrementally():
    """BatchedEnv counts only the instances whose topology changed (actions, maintenance, resets): after every step its cap equals the
    cap of a from-scratch count over all instances (host logic only: a stub stands where the engine's series calls go)."""
    from grid2op_b200.batched_env import BatchedEnv, random_substation_actions
    from grid2op_b200.engine import grid_max_active_buses, make_grid_desc
    from grid2op_b200.gridmodel import GridModel
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_neurips_2020_track1.npz"))
    B = 96

    class Stub:
        def __init__(self):
            self.desc, self.keep = make_grid_desc(gm)
            self.k = 0
            self.rows_counted = []

        def series_bind(self, *a): pass
        def series_set_topo(self, t): pass
        def series_step(self, **k): pass
        def series_next_is_reset(self): pass
        def series_reset_instances(self, *a, **k): pass

        def series_fetch(self, want_out=True, want_rho=True):
            self.k += 1
            st = (np.random.default_rng(self.k).random(B) < 0.15).astype(np.int32)
            return None, st, np.full(B, 4, np.int32), np.zeros((B, gm.n_line), np.float32)

        def max_active_buses(self, topo, per_instance=False):
            self.rows_counted.append(len(topo))
            r = grid_max_active_buses(gm, topo, self.desc)
            return r if per_instance else r[0]

    def scratch_cap(env):
        n = grid_max_active_buses(gm, env.topo)[0]
        return min((n + 7) // 8 * 8 if n > 17 else n, gm.n_slot)

    stub = Stub()
    chron = np.zeros((2, 50, 2 * gm.n_load + 2 * gm.n_gen), np.float32)
    env = BatchedEnv(gm, chron, B, engine=stub)
    assert env.nb_cap == scratch_cap(env)
    rng = np.random.default_rng(5)
    for step in range(25):
        if step % 3 == 2:                       # a step where only a few instances act
            sub, bus = random_substation_actions(env, rng)
            sub[8:] = -1
            stub.rows_counted.clear()
            _, done, _ = env.step(sub_id=sub, sub_bus=bus)
            assert max(stub.rows_counted, default=0) <= 8
        else:
            sub, bus = random_substation_actions(env, rng)
            _, done, _ = env.step(sub_id=sub, sub_bus=bus)
        assert env.nb_cap == scratch_cap(env), step
        d = np.flatnonzero(done)
        if len(d):
            stub.rows_counted.clear()
            env.reset_instances(d)
            assert stub.rows_counted == [len(d)]
            assert env.nb_cap == scratch_cap(env), step
