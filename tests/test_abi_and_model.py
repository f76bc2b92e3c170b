"""CPU tests: the C-ABI library loads and exports everything include/b200pf.h declares (no compute
calls without a GPU); the static grid model follows the reference's conventions."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REPO, env_grid

from grid2op_b200 import engine as eng_mod
from grid2op_b200.gridmodel import GridModel


def _declared_symbols():
    hdr = open(os.path.join(REPO, "include", "b200pf.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200pf_[a-z_0-9]+)\s*\(", hdr)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    assert set(syms) == set(eng_mod.ABI_SYMBOLS), (syms, eng_mod.ABI_SYMBOLS)


def test_library_loads_and_exports_every_declared_symbol():
    p = eng_mod.lib_path()
    assert os.path.exists(p), "libb200pf.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(p)
    for s in _declared_symbols():
        assert hasattr(lib, s), s
    lib.b200pf_abi_version.restype = ctypes.c_int
    assert lib.b200pf_abi_version() == 2


def test_engine_fails_loudly_without_a_gpu():
    """No CPU fallback: creating an engine without a CUDA device raises (on a GPU box it succeeds)."""
    path = env_grid("l2rpn_case14_sandbox")
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    lib = eng_mod.load_library()
    if lib.b200pf_device_count() > 0:
        e = eng_mod.PowerFlowEngine(gm, 1)
        e.close()
    else:
        with pytest.raises(eng_mod.EngineUnavailable):
            eng_mod.PowerFlowEngine(gm, 1)


EXPECT = {  # SURVEY.md section 8 size table (decoded from the reference's grid.json files)
    "rte_case5_example": dict(n_sub=5, n_line=8, n_gen=2, n_load=3, n_shunt=0, n_storage=0, dim_topo=21),
    "l2rpn_case14_sandbox": dict(n_sub=14, n_line=20, n_gen=6, n_load=11, n_shunt=1, n_storage=0, dim_topo=57),
    "l2rpn_neurips_2020_track1": dict(n_sub=36, n_line=59, n_gen=22, n_load=37, n_shunt=6, n_storage=0, dim_topo=177),
    "l2rpn_wcci_2022_dev": dict(n_sub=118, n_line=186, n_gen=62, n_load=91, n_shunt=14, n_storage=7, dim_topo=532),
}


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_grid_model_sizes_and_layout(name):
    path = env_grid(name)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    for k, v in EXPECT[name].items():
        assert getattr(gm, k) == v, k
    # every topo_vect position is used exactly once (reference GridObjects._compute_pos_big_topo_cls)
    pos = np.concatenate([gm.line_or_pos, gm.line_ex_pos, gm.gen_pos, gm.load_pos, gm.storage_pos])
    assert sorted(pos.tolist()) == list(range(gm.dim_topo))
    assert gm.sub_info.sum() == gm.dim_topo
    # branch admittances are finite and reciprocal for untapped lines
    assert np.isfinite(gm.line_y).all()
    nl = gm.n_powerline
    assert np.allclose(gm.line_y[:nl, 2:4], gm.line_y[:nl, 4:6])
    assert gm.n_out == 10 * gm.n_line + 4 * gm.n_unit + 2 * gm.n_load + gm.n_storage + 3 * gm.n_shunt


def test_names_follow_reference_conventions():
    path = env_grid("l2rpn_case14_sandbox")
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    # reference pPB:490-508: unnamed lines "{from}_{to}_{i}", trafos "{a}_{b}_{i+n_line}" with a STRING sort of (hv, lv)
    assert gm.name_line[0] == "0_1_0" and gm.name_line[15] == "3_6_15" and gm.name_line[-1] == "6_8_19"
    assert gm.name_gen[0] == "gen_1_0" and gm.name_load[0] == "load_1_0" and gm.name_shunt[0] == "shunt_8_0"
    # thermal limits pPB:806-813
    assert gm.thermal_limit_a.dtype == np.float32 and gm.thermal_limit_a.shape == (20,)


def test_reference_unit_normalisation_on_ext_grid_files():
    """Files without a slack generator get a generator on the ext_grid bus (reference pPB:394-453)."""
    path = env_grid("l2rpn_2019")
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    assert gm.n_hidden == 1 and gm.id_gen_added == gm.n_gen - 1 and gm.gen_slack[gm.id_gen_added]
    assert gm.unit_is_ref.tolist().count(1) == 2
