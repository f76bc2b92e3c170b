"""Backend.save_file (reference pPB:1425-1437): the written file is a pandapower-JSON grid that loads back with the
set points of the backend (host logic only: engine replaced by the oracle adapter)."""
import os

import numpy as np
import pytest

from conftest import env_grid


def test_save_file_roundtrip(tmp_path):
    path = env_grid("l2rpn_case14_sandbox")
    if path is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    from grid2op_b200.gridmodel import GridModel
    from oracle_engine import OracleEngine

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    b = HostLogicBackend()
    b.load_grid(path)
    b._load_p[:] *= 1.1
    b._gen_vm[0] = 1.02
    b._line_on[3] = False
    out = os.path.join(tmp_path, "saved.json")
    b.save_file(out)
    gm2 = GridModel(out)
    assert np.allclose(gm2.load_p0, b._load_p) and np.isclose(gm2.gen_vm0[0], 1.02)
    assert not gm2.line_in_service0[3] and gm2.line_in_service0[[0, 1, 2, 4]].all()
    assert gm2.n_line == b._gm.n_line and np.array_equal(gm2.line_y, b._gm.line_y)


def test_saved_file_is_a_complete_net_of_the_current_topology(tmp_path):
    """with elements on busbar 2 and a line out: the saved file carries the duplicated busbar buses like the reference's
    ``pp.to_json`` (pPB:548-566, 1425-1437), every bus id an element names exists, and a power flow of the FILE (the oracle's
    pandapower restatement: from_json + runpp) gives the flows the backend reports"""
    path = env_grid("l2rpn_case14_sandbox")
    if path is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    from grid2op_b200.ppjson import read_pp_json
    from oracle import pandapower_ref as ppr
    from oracle_engine import OracleEngine

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    b = HostLogicBackend()
    b.load_grid(path)
    gm = b._gm
    # substation 1: two line ends + the load to busbar 2 (a valid split), line 7 out
    on_sub1 = [l for l in range(gm.n_line) if gm.line_or_sub[l] == 1][:2]
    for l in on_sub1:
        b._lor_bus[l] = 2
    ex_sub1 = [l for l in range(gm.n_line) if gm.line_ex_sub[l] == 1][:1]
    for l in ex_sub1:
        b._lex_bus[l] = 2
    b._line_on[7] = False
    b._load_p[:] *= 1.05
    conv, exc = b.runpf()
    assert conv, exc
    p_or = b.lines_or_info()[0].copy()
    out = os.path.join(tmp_path, "saved_split.json")
    b.save_file(out)
    net = read_pp_json(out)
    bus = net.table("bus")
    labels = set(int(v) for v in bus.index)
    assert labels == set(range(2 * gm.n_sub))
    on = dict(zip((int(v) for v in bus.index), bus.col("in_service", 1.0) != 0))
    assert on[1] and on[1 + gm.n_sub] and not on[0 + gm.n_sub]
    for tname, cols in (("line", ("from_bus", "to_bus")), ("trafo", ("hv_bus", "lv_bus")), ("load", ("bus",)), ("gen", ("bus",)), ("shunt", ("bus",))):
        t = net.table(tname)
        for c in cols:
            if len(t):
                assert set(int(v) for v in t.col(c, 0)) <= labels
    assert sorted(int(v) for v in net.table("line").col("from_bus", 0)[on_sub1]) == [1 + gm.n_sub] * len(on_sub1)
    # the file, solved as a pandapower net
    pnet = ppr.from_json(out)
    ppr.runpp(pnet)
    assert pnet.converged
    res = np.concatenate([np.asarray(pnet.res["line"]["p_from_mw"], dtype=float), np.asarray(pnet.res["trafo"]["p_hv_mw"], dtype=float)])
    live = np.flatnonzero(b._line_on)
    assert np.allclose(res[live], p_or[live], atol=2e-3), float(np.max(np.abs(res[live] - p_or[live])))
    assert abs(p_or[7]) == 0.0
