"""Reader for the pandapower-JSON wire format of ``grid.json`` files.

The reference loads these files with ``pp.from_json`` (reference:
grid2op/Backend/pandaPowerBackend.py:377).  pandapower is a third-party package that is
not part of this build, so this module restates just the *container* format: a top level
``{"_module", "_class": "pandapowerNet", "_object": ...}`` whose ``_object`` is either a
dict of tables (pandapower >= 2.2 files) or a JSON *string* of that dict (older files, e.g.
grid2op/data_test/test_PandaPower/test_case14.json).  Each table is
``{"_class": "DataFrame", "orient": "split"|"columns", "dtype": {...}, "_object": "<json>"}``.

Rows are kept in **file order** (the reference indexes elements positionally with
``enumerate(df.iterrows())``, pandaPowerBackend.py:491-553, 691-751), and the ``index``
labels are kept beside them because bus ids in element tables are bus *labels*.

No pandas involved: a table is a :class:`Table` = ordered dict ``column -> numpy array``.
"""
from __future__ import annotations

import json
from collections import OrderedDict
from typing import Any, Dict, Iterable, List, Optional

import numpy as np

__all__ = ["Table", "read_pp_json", "PPNet", "update_pp_json"]

_FLOAT_KINDS = ("float", "int", "uint", "bool")


class Table:
    """A minimal column store: ``index`` (labels, file order) + named numpy columns."""

    def __init__(self, index: np.ndarray, columns: "OrderedDict[str, np.ndarray]"):
        self.index = index
        self.columns = columns

    def __len__(self) -> int:
        return int(self.index.shape[0])

    def __contains__(self, name: str) -> bool:
        return name in self.columns

    def __getitem__(self, name: str) -> np.ndarray:
        return self.columns[name]

    def get(self, name: str, default=None):
        return self.columns.get(name, default)

    def col(self, name: str, default: float, dtype=np.float64) -> np.ndarray:
        """Column as a numeric array, missing column / null entries -> ``default``."""
        n = len(self)
        if name not in self.columns:
            return np.full(n, default, dtype=dtype)
        raw = self.columns[name]
        out = np.empty(n, dtype=np.float64)
        for i, v in enumerate(raw):
            if v is None:
                out[i] = np.nan
            elif isinstance(v, str):
                try:
                    out[i] = float(v)
                except ValueError:
                    out[i] = np.nan
            else:
                out[i] = float(v)
        out[np.isnan(out)] = default
        return out.astype(dtype)

    def strings(self, name: str) -> List[Optional[str]]:
        if name not in self.columns:
            return [None] * len(self)
        return [None if v is None else v for v in self.columns[name]]

    def has_nulls(self, name: str) -> bool:
        if name not in self.columns:
            return True
        for v in self.columns[name]:
            if v is None:
                return True
            if isinstance(v, float) and np.isnan(v):
                return True
        return False


def _decode_table(spec: Dict[str, Any]) -> Table:
    obj = spec["_object"]
    if isinstance(obj, str):
        obj = json.loads(obj)
    orient = spec.get("orient", "split")
    dtypes = spec.get("dtype", {}) or {}
    cols: "OrderedDict[str, np.ndarray]" = OrderedDict()
    if orient == "split" or ("columns" in obj and "data" in obj and "index" in obj):
        names = list(obj["columns"])
        index = list(obj["index"])
        data = obj["data"]
        for j, nm in enumerate(names):
            cols[str(nm)] = np.array([row[j] for row in data], dtype=object)
    else:
        # orient == "columns": {col: {index_label: value}}
        names = list(obj.keys())
        index = None
        for nm in names:
            colmap = obj[nm]
            if index is None:
                index = list(colmap.keys())
            cols[str(nm)] = np.array([colmap.get(k) for k in index], dtype=object)
        if index is None:
            index = []
    idx_arr = np.empty(len(index), dtype=np.int64)
    for i, v in enumerate(index):
        try:
            idx_arr[i] = int(v)
        except (TypeError, ValueError):
            idx_arr[i] = i
    # honour the dtype map for object columns: keep strings as they are (names such as
    # "80_79_175" must not become numbers)
    for nm, dt in dtypes.items():
        if nm in cols and str(dt) == "object":
            cols[nm] = np.array([None if v is None else (v if not isinstance(v, (int, float)) or isinstance(v, bool) else v)
                                 for v in cols[nm]], dtype=object)
    return Table(idx_arr, cols)


class PPNet:
    """The decoded network: scalar fields + :class:`Table` per element kind."""

    def __init__(self, fields: Dict[str, Any], tables: Dict[str, Table]):
        self.fields = fields
        self.tables = tables

    def table(self, name: str) -> Table:
        if name in self.tables:
            return self.tables[name]
        return Table(np.zeros(0, dtype=np.int64), OrderedDict())

    @property
    def sn_mva(self) -> float:
        return float(self.fields.get("sn_mva", 1.0))

    @property
    def f_hz(self) -> float:
        return float(self.fields.get("f_hz", 50.0))

    @property
    def version(self) -> str:
        return str(self.fields.get("version", ""))


_ELEMENT_TABLES = ("bus", "load", "sgen", "storage", "gen", "shunt", "ext_grid", "line", "trafo",
                   "trafo3w", "impedance", "dcline", "ward", "xward", "switch", "motor",
                   "asymmetric_load", "asymmetric_sgen", "measurement")


def read_pp_json(path: str, tables: Iterable[str] = _ELEMENT_TABLES) -> PPNet:
    with open(path, "r", encoding="utf-8") as f:
        top = json.load(f)
    obj = top["_object"] if isinstance(top, dict) and "_object" in top else top
    if isinstance(obj, str):
        obj = json.loads(obj)
    fields: Dict[str, Any] = {}
    out: Dict[str, Table] = {}
    for k, v in obj.items():
        if isinstance(v, dict) and v.get("_class") == "DataFrame":
            if k in tables:
                out[k] = _decode_table(v)
        elif not isinstance(v, (dict, list)):
            fields[k] = v
    return PPNet(fields, out)


def update_pp_json(src_path: str, dst_path: str, updates: Dict[str, Dict[str, Any]],
                   duplicate_rows: Optional[Dict[str, List[Any]]] = None) -> None:
    """Write a copy of the pandapower-JSON file ``src_path`` to ``dst_path`` with some table columns replaced
    (``updates[table][column] = sequence of new values, file row order``) - the part of ``pp.to_json`` that
    ``Backend.save_file`` needs (reference: grid2op/Backend/pandaPowerBackend.py:1425-1437).  Everything else
    (other tables, std_types, geodata, options) is carried over byte-for-byte; result tables are emptied of
    nothing - they simply keep the values stored in the source file.

    ``duplicate_rows[table] = [(source row number, new index label), ...]``: copies of existing rows appended to the table
    BEFORE the column updates (``updates`` then names values for the old and the appended rows) — the busbar buses the
    reference creates at load time (pPB:548-566) and therefore writes with ``pp.to_json``."""
    with open(src_path, "r", encoding="utf-8") as f:
        top = json.load(f)
    obj = top["_object"] if isinstance(top, dict) and "_object" in top else top
    as_string = isinstance(obj, str)
    if as_string:
        obj = json.loads(obj)
    for tname, cols in updates.items():
        spec = obj.get(tname)
        if not (isinstance(spec, dict) and spec.get("_class") == "DataFrame"):
            raise KeyError(f"table {tname!r} not found in {src_path}")
        inner = spec["_object"]
        inner_is_str = isinstance(inner, str)
        tab = json.loads(inner) if inner_is_str else inner
        if not ("columns" in tab and "data" in tab):
            # orient="columns" -> convert to split so that rows can be edited uniformly
            index = None
            names = list(tab.keys())
            for nm in names:
                if index is None:
                    index = list(tab[nm].keys())
            data = [[tab[nm].get(k) for nm in names] for k in (index or [])]
            tab = {"columns": names, "index": index or [], "data": data}
            spec["orient"] = "split"
        for src_row, label in (duplicate_rows or {}).get(tname, []):
            tab["data"].append(list(tab["data"][int(src_row)]))
            if "index" in tab and tab["index"] is not None:
                tab["index"].append(int(label))
        for cname, values in cols.items():
            values = list(values)
            if cname not in tab["columns"]:
                tab["columns"].append(cname)
                for row in tab["data"]:
                    row.append(None)
            j = tab["columns"].index(cname)
            if len(values) != len(tab["data"]):
                raise ValueError(f"{tname}.{cname}: {len(values)} values for {len(tab['data'])} rows")
            for row, v in zip(tab["data"], values):
                if isinstance(v, (np.bool_, bool)):
                    row[j] = bool(v)
                elif isinstance(v, (np.integer,)):
                    row[j] = int(v)
                elif isinstance(v, (np.floating, float)):
                    row[j] = float(v)
                else:
                    row[j] = v
        spec["_object"] = json.dumps(tab) if inner_is_str else tab
    if isinstance(top, dict) and "_object" in top:
        top["_object"] = json.dumps(obj) if as_string else obj
    else:
        top = obj
    with open(dst_path, "w", encoding="utf-8") as f:
        json.dump(top, f)
