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
