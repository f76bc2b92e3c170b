"""The C restatement (oracle/pf_oracle.c, used as CPU baseline) agrees with the numpy oracle, on the file
state of every grid and on random topologies / injections, AC and DC, including diverging cases."""
import numpy as np
import pytest

from conftest import env_grid

from grid2op_b200.gridmodel import GridModel
from oracle.c_oracle import COracle
from oracle_engine import OracleEngine

GRIDS = ["rte_case5_example", "l2rpn_case14_sandbox", "educ_case14_storage", "l2rpn_neurips_2020_track1", "l2rpn_2019"]


def random_cases(gm, n, seed, p_split=0.15, p_disc=0.03):
    rng = np.random.default_rng(seed)
    topo = np.tile(gm.default_topo(), (n, 1))
    inj = np.tile(gm.default_inj(), (n, 1))
    sl = gm.inj_slices()
    for i in range(n):
        if i == 0:
            continue
        # move whole groups of elements of a few substations to busbar 2, disconnect a few lines
        for s in rng.choice(gm.n_sub, size=rng.integers(0, 3), replace=False):
            pos = [p for p in range(gm.dim_topo) if _sub_of(gm, p) == s]
            for p in pos:
                if topo[i, p] > 0 and rng.random() < 0.5:
                    topo[i, p] = 2
        for l in range(gm.n_line):
            if rng.random() < p_disc:
                topo[i, gm.line_or_pos[l]] = -1
                topo[i, gm.line_ex_pos[l]] = -1
        inj[i, sl["load_p"]] *= rng.uniform(0.8, 1.2, gm.n_load)
        inj[i, sl["load_q"]] *= rng.uniform(0.8, 1.2, gm.n_load)
        inj[i, sl["gen_p"]] *= rng.uniform(0.8, 1.2, gm.n_gen)
    return topo, inj


def _sub_of(gm, p):
    m = getattr(gm, "_test_sub_of_pos", None)
    if m is None:
        m = np.zeros(gm.dim_topo, dtype=np.int64)
        m[gm.line_or_pos] = gm.line_or_sub; m[gm.line_ex_pos] = gm.line_ex_sub
        m[gm.gen_pos] = gm.gen_sub; m[gm.load_pos] = gm.load_sub
        if gm.n_storage:
            m[gm.storage_pos] = gm.storage_sub
        gm._test_sub_of_pos = m
    return m[p]


@pytest.mark.parametrize("name", GRIDS)
@pytest.mark.parametrize("dc", [False, True])
def test_c_oracle_matches_numpy_oracle(name, dc):
    path = env_grid(name)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    topo, inj = random_cases(gm, 24, seed=7)
    o1, s1, i1, _ = OracleEngine(gm).run(topo, inj, is_dc=dc)
    o2, s2, i2, _ = COracle(gm, nthreads=2).run(topo, inj, is_dc=dc)
    assert (s1 == 0).sum() >= 8, "want a healthy share of converging cases"
    assert np.array_equal(s1 == 0, s2 == 0), (s1, s2)
    ok = s1 == 0
    scale = 1.0 + np.abs(o1[ok])
    assert np.max(np.abs(o1[ok] - o2[ok]) / scale) <= 2e-5     # both fp64 inside; float32 records
    if not dc:
        assert np.array_equal(i1[ok], i2[ok])
