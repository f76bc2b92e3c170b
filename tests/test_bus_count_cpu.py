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
