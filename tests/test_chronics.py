"""The package's chronics reader returns what grid2op's own GridStateFromFile serves to the backend
(reference grid2op/Chronics/gridStateFromFile.py:749-808): same rows, columns re-ordered by NAME."""
import os
import warnings

import numpy as np
import pytest

from conftest import env_grid, grid2op_root

from grid2op_b200.chronics import load_scenario, load_scenarios
from grid2op_b200.gridmodel import GridModel


@pytest.mark.parametrize("env_name,scen", [("l2rpn_case14_sandbox", "0001"), ("rte_case5_example", "03")])
def test_reader_matches_grid2op_loader(env_name, scen):
    path = env_grid(env_name)
    if path is None:
        pytest.skip("reference data not available")
    from grid2op_b200._bootstrap import ensure_grid2op
    if not ensure_grid2op():
        pytest.skip("grid2op not importable")
    from grid2op.Chronics import GridStateFromFile
    gm = GridModel(path)
    folder = os.path.join(os.path.dirname(path), "chronics", scen)
    mine = load_scenario(folder, gm)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gs = GridStateFromFile(path=folder, sep=";")
        gs.initialize(order_backend_loads=list(gm.name_load), order_backend_prods=list(gm.name_gen),
                      order_backend_lines=list(gm.name_line), order_backend_subs=list(gm.name_sub))
    nl, ng = gm.n_load, gm.n_gen
    n = min(mine.shape[0], gs.load_p.shape[0])
    assert n >= 100
    assert np.array_equal(mine[:n, :nl], gs.load_p[:n].astype(np.float32))
    assert np.array_equal(mine[:n, nl:2 * nl], gs.load_q[:n].astype(np.float32))
    assert np.array_equal(mine[:n, 2 * nl:2 * nl + ng], gs.prod_p[:n].astype(np.float32))
    assert np.array_equal(mine[:n, 2 * nl + ng:], gs.prod_v[:n].astype(np.float32))


def test_committed_fixture_is_what_the_reader_produces():
    path = env_grid("l2rpn_case14_sandbox")
    if path is None:
        pytest.skip("reference data not available")
    gm = GridModel(path)
    chron = load_scenarios(os.path.join(os.path.dirname(path), "chronics"), gm)
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", "case14_sandbox_chronics.npz"))["chron"]
    assert np.array_equal(chron, fix)


@pytest.mark.parametrize("env_name,scen", [("rte_case5_example", "00"), ("rte_case5_example", "07"), ("l2rpn_case14_sandbox", "0000")])
def test_forecasts_and_outage_tables_match_grid2op_loader(env_name, scen):
    """*_forecasted tables, maintenance / hazards and their derived series == what GridStateFromFileWithForecasts holds
    (reference grid2op/Chronics/gridStateFromFileWithForecasts.py:110-190, gridStateFromFile.py:645-682)."""
    path = env_grid(env_name)
    if path is None:
        pytest.skip("reference data not available")
    from grid2op_b200._bootstrap import ensure_grid2op
    if not ensure_grid2op():
        pytest.skip("grid2op not importable")
    from grid2op.Chronics import GridStateFromFileWithForecasts
    from grid2op_b200.chronics import load_forecasts, load_line_events, maintenance_time_duration, hazard_duration
    gm = GridModel(path)
    folder = os.path.join(os.path.dirname(path), "chronics", scen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gs = GridStateFromFileWithForecasts(path=folder, sep=";")
        gs.initialize(order_backend_loads=list(gm.name_load), order_backend_prods=list(gm.name_gen),
                      order_backend_lines=list(gm.name_line), order_backend_subs=list(gm.name_sub))
    fc = load_forecasts(folder, gm)
    assert fc is not None
    nl, ng = gm.n_load, gm.n_gen
    n = min(fc.shape[0], gs.load_p_forecast.shape[0])
    assert n >= 100
    assert np.array_equal(fc[:n, :nl], gs.load_p_forecast[:n].astype(np.float32))
    assert np.array_equal(fc[:n, nl:2 * nl], gs.load_q_forecast[:n].astype(np.float32))
    assert np.array_equal(fc[:n, 2 * nl:2 * nl + ng], gs.prod_p_forecast[:n].astype(np.float32))
    assert np.array_equal(fc[:n, 2 * nl + ng:], gs.prod_v_forecast[:n].astype(np.float32))
    maint = load_line_events(folder, gm, "maintenance")
    haz = load_line_events(folder, gm, "hazards")
    if gs.maintenance is None:
        assert maint is None
    else:
        m = min(len(maint), len(gs.maintenance))
        assert np.array_equal(maint[:m], gs.maintenance[:m])
        t, d = maintenance_time_duration(maint)
        assert np.array_equal(t[:m], gs.maintenance_time[:m]) and np.array_equal(d[:m], gs.maintenance_duration[:m])
    if gs.hazards is None:
        assert haz is None
    else:
        m = min(len(haz), len(gs.hazards))
        assert np.array_equal(haz[:m], gs.hazards[:m])
        assert np.array_equal(hazard_duration(haz)[:m], gs.hazard_duration[:m])


def test_maintenance_and_hazard_series_match_gridvalue_helpers():
    """Derived series on random outage patterns == the reference's own helpers
    (grid2op/Chronics/gridValue.py:264-470).  Patterns start with a free step: a planned outage already running at row 0
    is outside what the reference's diff-based helpers handle."""
    from grid2op_b200._bootstrap import ensure_grid2op
    if not ensure_grid2op():
        pytest.skip("grid2op not importable")
    from grid2op.Chronics import GridValue
    from grid2op_b200.chronics import maintenance_time_duration, hazard_duration
    rng = np.random.default_rng(3)
    m = rng.random((400, 12)) < 0.08
    for _ in range(3):                       # make runs longer than one step
        m[1:] |= m[:-1] & (rng.random((399, 12)) < 0.6)
    m[0] = False
    t, d = maintenance_time_duration(m)
    hd = hazard_duration(m)
    for l in range(m.shape[1]):
        col = m[:, l].astype(np.int32)
        assert np.array_equal(t[:, l], GridValue.get_maintenance_time_1d(col))
        assert np.array_equal(d[:, l], GridValue.get_maintenance_duration_1d(col))
        assert np.array_equal(hd[:, l], GridValue.get_hazard_duration_1d(col))
    assert m.any() and (t > 0).any() and (d > 1).any()
