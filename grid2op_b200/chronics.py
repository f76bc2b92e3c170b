"""Chronics ingest: grid2op time-series folders (csv / csv.bz2, ';' separated) -> one float32 array in
BACKEND element order, ready to be pinned and copied to HBM (``PowerFlowEngine.series_bind``).

Restates the part of ``grid2op.Chronics.GridStateFromFile`` the batched driver needs (reference
grid2op/Chronics/gridStateFromFile.py:289-387 file discovery and column matching by NAME, :749-808 one row
per step): ``load_p, load_q, prod_p, prod_v``; a missing ``prod_v`` file falls back to the set points of
the grid file (the reference keeps the backend's values in that case).
Also read: the ``*_forecasted`` tables of ``GridStateFromFileWithForecasts``
(grid2op/Chronics/gridStateFromFileWithForecasts.py:110-190; what ``obs.simulate`` is fed with) and the planned /
unplanned outage tables ``maintenance`` / ``hazards`` with the derived ``maintenance_time``,
``maintenance_duration``, ``hazard_duration`` series (grid2op/Chronics/gridValue.py:264-470), columns matched by
LINE name.
"""
from __future__ import annotations

import bz2
import os
from typing import List, Optional, Sequence

import numpy as np

from .gridmodel import GridModel

__all__ = ["read_table", "load_scenario", "load_scenarios", "list_scenarios", "load_forecasts", "load_line_events",
           "maintenance_time_duration", "hazard_duration"]


def _open(path_no_ext: str):
    for ext, opener in ((".csv.bz2", lambda p: bz2.open(p, "rt")), (".csv", lambda p: open(p, "rt"))):
        p = path_no_ext + ext
        if os.path.exists(p):
            return opener(p)
    return None


def read_table(path_no_ext: str, sep: str = ";"):
    """-> (column names, float64 array [n_rows, n_cols]) or None when the file does not exist."""
    f = _open(path_no_ext)
    if f is None:
        return None
    with f:
        header = f.readline().strip().split(sep)
        rows = [line.split(sep) for line in f if line.strip()]
    return header, np.array(rows, dtype=np.float64)


def _ordered(tbl, names: Sequence[str], what: str, folder: str) -> np.ndarray:
    header, arr = tbl
    pos = {n: i for i, n in enumerate(header)}
    missing = [n for n in names if n not in pos]
    if missing:
        raise ValueError(f"{what} of {folder}: no column for {missing} (columns are matched by name, "
                         "like grid2op GridStateFromFile)")
    return arr[:, [pos[n] for n in names]]


def load_scenario(folder: str, gm: GridModel, sep: str = ";", suffix: str = "") -> np.ndarray:
    """float32 [n_rows, 2 n_load + 2 n_gen] = load_p | load_q | prod_p | prod_v[kV], backend order."""
    lp = read_table(os.path.join(folder, "load_p" + suffix), sep)
    lq = read_table(os.path.join(folder, "load_q" + suffix), sep)
    pp = read_table(os.path.join(folder, "prod_p" + suffix), sep)
    pv = read_table(os.path.join(folder, "prod_v" + suffix), sep)
    if lp is None or lq is None or pp is None:
        raise FileNotFoundError(f"{folder}: load_p{suffix} / load_q{suffix} / prod_p{suffix} are required")
    a_lp = _ordered(lp, gm.name_load, "load_p", folder)
    a_lq = _ordered(lq, gm.name_load, "load_q", folder)
    a_pp = _ordered(pp, gm.name_gen, "prod_p", folder)
    n = min(len(a_lp), len(a_lq), len(a_pp))
    if pv is not None:
        a_pv = _ordered(pv, gm.name_gen, "prod_v", folder)
        n = min(n, len(a_pv))
    else:
        a_pv = np.tile((gm.gen_vm0 * gm.prod_pu_to_kv.astype(np.float64))[None, :], (n, 1))
    return np.concatenate([a_lp[:n], a_lq[:n], a_pp[:n], a_pv[:n]], axis=1).astype(np.float32)


def load_forecasts(folder: str, gm: GridModel, sep: str = ";") -> Optional[np.ndarray]:
    """The ``*_forecasted`` tables in the layout of :func:`load_scenario` (row k = forecast available at step k for
    step k + 1, one horizon like the bundled environments), or None when the folder has no forecasts.  These rows feed
    ``PowerFlowEngine.run`` / the rows entry points exactly like real rows: a batched ``obs.simulate`` is one launch."""
    if _open(os.path.join(folder, "load_p_forecasted")) is None:
        return None
    return load_scenario(folder, gm, sep, suffix="_forecasted")


def load_line_events(folder: str, gm: GridModel, what: str = "maintenance", sep: str = ";") -> Optional[np.ndarray]:
    """bool [n_rows, n_line] from ``maintenance`` / ``hazards`` (non-zero = the line is out), backend line order,
    columns matched by line name; None when the file does not exist."""
    tbl = read_table(os.path.join(folder, what), sep)
    if tbl is None:
        return None
    return np.abs(_ordered(tbl, gm.name_line, what, folder)) >= 1e-7


def _runs(col: np.ndarray):
    """(start, end) index pairs of the runs of True in a 1-d bool array."""
    x = np.concatenate(([0], col.astype(np.int8), [0]))
    d = np.diff(x)
    return np.flatnonzero(d == 1), np.flatnonzero(d == -1)


def maintenance_time_duration(maintenance: np.ndarray):
    """int32 [n_rows, n_line] x 2: steps until the next planned outage starts (0 while it runs, -1 when none is
    ahead) and its (remaining) duration (0 when none is ahead) — the semantics documented at
    grid2op/Chronics/gridValue.py:264-395."""
    n, nl = maintenance.shape
    t = np.full((n, nl), -1, dtype=np.int32)
    dur = np.zeros((n, nl), dtype=np.int32)
    for l in range(nl):
        prev = 0
        for b, e in zip(*_runs(maintenance[:, l])):
            t[prev:b, l] = np.arange(b - prev, 0, -1)
            t[b:e, l] = 0
            dur[prev:b, l] = e - b
            dur[b:e, l] = np.arange(e - b, 0, -1)
            prev = e
    return t, dur


def hazard_duration(hazards: np.ndarray) -> np.ndarray:
    """int32 [n_rows, n_line]: remaining duration of the running unplanned outage, 0 otherwise
    (grid2op/Chronics/gridValue.py:403-470)."""
    n, nl = hazards.shape
    dur = np.zeros((n, nl), dtype=np.int32)
    for l in range(nl):
        for b, e in zip(*_runs(hazards[:, l])):
            dur[b:e, l] = np.arange(e - b, 0, -1)
    return dur


def list_scenarios(chronics_dir: str) -> List[str]:
    return sorted(os.path.join(chronics_dir, d) for d in os.listdir(chronics_dir)
                  if os.path.isdir(os.path.join(chronics_dir, d)))


def load_scenarios(chronics_dir: str, gm: GridModel, scenarios: Optional[Sequence[str]] = None, sep: str = ";") -> np.ndarray:
    """float32 [n_scen, n_rows, ncol]; scenarios of different length are padded by repeating their last row
    (``lengths`` of the originals are returned as the ``.lengths`` attribute of a companion array)."""
    folders = list(scenarios) if scenarios is not None else list_scenarios(chronics_dir)
    arrs = [load_scenario(f, gm, sep) for f in folders]
    n = max(a.shape[0] for a in arrs)
    out = np.empty((len(arrs), n, arrs[0].shape[1]), dtype=np.float32)
    for i, a in enumerate(arrs):
        out[i, :a.shape[0]] = a
        out[i, a.shape[0]:] = a[-1]
    return out
