"""Chronics ingest: grid2op time-series folders (csv / csv.bz2, ';' separated) -> one float32 array in
BACKEND element order, ready to be pinned and copied to HBM (``PowerFlowEngine.series_bind``).

Restates the part of ``grid2op.Chronics.GridStateFromFile`` the batched driver needs (reference
grid2op/Chronics/gridStateFromFile.py:289-387 file discovery and column matching by NAME, :749-808 one row
per step): ``load_p, load_q, prod_p, prod_v``; a missing ``prod_v`` file falls back to the set points of
the grid file (the reference keeps the backend's values in that case).  Maintenance / hazards / forecasts
are not read here (DoNothing driver).
"""
from __future__ import annotations

import bz2
import os
from typing import List, Optional, Sequence

import numpy as np

from .gridmodel import GridModel

__all__ = ["read_table", "load_scenario", "load_scenarios", "list_scenarios"]


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


def load_scenario(folder: str, gm: GridModel, sep: str = ";") -> np.ndarray:
    """float32 [n_rows, 2 n_load + 2 n_gen] = load_p | load_q | prod_p | prod_v[kV], backend order."""
    lp = read_table(os.path.join(folder, "load_p"), sep)
    lq = read_table(os.path.join(folder, "load_q"), sep)
    pp = read_table(os.path.join(folder, "prod_p"), sep)
    pv = read_table(os.path.join(folder, "prod_v"), sep)
    if lp is None or lq is None or pp is None:
        raise FileNotFoundError(f"{folder}: load_p / load_q / prod_p are required")
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
