"""ORACLE (test infrastructure, never shipped / never on the product path).

CPU fp64 restatement of the arithmetic that ``PandaPowerBackend.runpf`` delegates to the
third-party package **pandapower** (``pandapower>=3.1.1``, reference ``pyproject.toml:15``;
the bundled grids were saved by pandapower 2.8.0 / 2.0.1).  pandapower is *not* under
``/root/reference`` and is not installed here, so its published algorithm is restated from
its documented element model and the PYPOWER/MATPOWER formulation it is built on:

* ``pp.from_json``                (call site: grid2op/Backend/pandaPowerBackend.py:377)
* ``pp.runpp(check_connectivity=False, init="dc", max_iteration=10, distributed_slack=False)``
                                  (call site: pandaPowerBackend.py:1097-1105)
* ``pp.rundcpp(check_connectivity=True, init="flat")``   (call site: pandaPowerBackend.py:1090)
* the ``res_*`` tables read back by pandaPowerBackend.py:1122-1218, 1526-1647.

Pinned by (see tests/test_oracle_golden.py):
  - the ``res_bus/res_line/res_trafo/res_gen/res_shunt`` tables stored *inside* the
    reference's own ``grid.json`` files by real pandapower (rte_case5_example,
    l2rpn_neurips_2020_track1, l2rpn_wcci_2022_dev, rte_case118_example, test_case14.json),
  - grid2op/tests/BaseBackendTest.py:258-319, 438-530 golden p_or/q_or vectors,
  - grid2op/data/rte_case5_example/_statistics/*.npz (7930 recorded AC steps).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this.
"""
from __future__ import annotations

import copy
import json
from typing import Dict, List, Optional

import numpy as np

SQRT3 = np.sqrt(3.0)


class LoadflowNotConverged(Exception):
    """Mirror of ``pandapower.powerflow.LoadflowNotConverged`` (caught at pandaPowerBackend.py:1250)."""


# --------------------------------------------------------------------------------------
# container: a "DataFrame" is a dict of equally long numpy columns + "index" labels
# --------------------------------------------------------------------------------------
class Frame:
    def __init__(self, index, cols: Dict[str, np.ndarray]):
        self.index = np.asarray(index, dtype=np.int64)
        self.cols = cols

    def __len__(self):
        return int(self.index.shape[0])

    def __contains__(self, k):
        return k in self.cols

    def __getitem__(self, k):
        return self.cols[k]

    def __setitem__(self, k, v):
        self.cols[k] = v

    def num(self, k, default):
        """float64 view of a column; missing column / nulls -> default."""
        n = len(self)
        if k not in self.cols:
            return np.full(n, float(default))
        out = np.empty(n, dtype=np.float64)
        for i, v in enumerate(self.cols[k]):
            if v is None or isinstance(v, str):
                out[i] = np.nan
            else:
                out[i] = float(v)
        out[np.isnan(out)] = default
        return out

    def flag(self, k, default=True):
        n = len(self)
        if k not in self.cols:
            return np.full(n, bool(default))
        return np.array([bool(v) if v is not None else bool(default) for v in self.cols[k]], dtype=bool)

    def text(self, k):
        if k not in self.cols:
            return [None] * len(self)
        return list(self.cols[k])

    def append_row(self, label: int, **values):
        """``pp.create_bus`` / ``pp.create_gen`` style append (missing columns -> None)."""
        self.index = np.concatenate([self.index, [label]])
        for k in list(self.cols.keys()):
            v = values.get(k, None)
            self.cols[k] = np.concatenate([np.asarray(self.cols[k], dtype=object), np.array([v], dtype=object)])
        for k, v in values.items():
            if k not in self.cols:
                col = np.empty(len(self), dtype=object)
                col[:] = None
                col[-1] = v
                self.cols[k] = col

    def copy(self):
        return Frame(self.index.copy(), {k: np.array(v, dtype=object).copy() for k, v in self.cols.items()})


def _decode_frame(spec) -> Frame:
    obj = spec["_object"]
    if isinstance(obj, str):
        obj = json.loads(obj)
    if "columns" in obj and "data" in obj:          # orient="split"
        names, index, data = obj["columns"], obj["index"], obj["data"]
        cols = {str(nm): np.array([r[j] for r in data], dtype=object) for j, nm in enumerate(names)}
    else:                                           # orient="columns"
        index = None
        cols = {}
        for nm, cm in obj.items():
            if index is None:
                index = list(cm.keys())
            cols[str(nm)] = np.array([cm.get(k) for k in index], dtype=object)
        index = index or []
    lab = []
    for i, v in enumerate(index):
        try:
            lab.append(int(v))
        except (TypeError, ValueError):
            lab.append(i)
    return Frame(lab, cols)


_EMPTY_COLS = {
    "bus": ["name", "vn_kv", "type", "zone", "in_service"],
    "line": ["name", "from_bus", "to_bus", "length_km", "r_ohm_per_km", "x_ohm_per_km", "c_nf_per_km",
             "g_us_per_km", "max_i_ka", "df", "parallel", "in_service"],
    "trafo": ["name", "hv_bus", "lv_bus", "sn_mva", "vn_hv_kv", "vn_lv_kv", "vk_percent", "vkr_percent",
              "pfe_kw", "i0_percent", "shift_degree", "tap_side", "tap_neutral", "tap_step_percent",
              "tap_pos", "parallel", "in_service"],
    "gen": ["name", "bus", "p_mw", "vm_pu", "min_q_mvar", "max_q_mvar", "scaling", "slack", "in_service"],
    "ext_grid": ["name", "bus", "vm_pu", "va_degree", "in_service"],
    "load": ["name", "bus", "p_mw", "q_mvar", "scaling", "in_service"],
    "shunt": ["name", "bus", "q_mvar", "p_mw", "vn_kv", "step", "in_service"],
    "storage": ["name", "bus", "p_mw", "q_mvar", "scaling", "in_service"],
    "sgen": ["name", "bus", "p_mw", "q_mvar", "scaling", "in_service"],
}


class Net:
    """Stand-in for ``pandapowerNet``: element tables + scalars + (after a run) ``res_*``."""

    def __init__(self):
        self.sn_mva = 1.0
        self.f_hz = 50.0
        self.version = ""
        self.converged = False
        self.tables: Dict[str, Frame] = {}
        self.stored_res: Dict[str, Frame] = {}   # res_* tables found in the file (goldens)
        self.res: Dict[str, Dict[str, np.ndarray]] = {}
        self.ppc_internal: Dict[str, np.ndarray] = {}

    def __getattr__(self, item):
        tables = self.__dict__.get("tables", {})
        if item in tables:
            return tables[item]
        raise AttributeError(item)

    def deepcopy(self):
        return copy.deepcopy(self)


def from_json(path: str) -> Net:
    """Restates the container side of ``pp.from_json`` (pandaPowerBackend.py:377).  Row order is
    file order; index labels kept; both on-disk dialects (dict ``_object`` / JSON-string ``_object``)."""
    with open(path, "r", encoding="utf-8") as f:
        top = json.load(f)
    obj = top["_object"] if isinstance(top, dict) and "_object" in top else top
    if isinstance(obj, str):
        obj = json.loads(obj)
    net = Net()
    for k, v in obj.items():
        if isinstance(v, dict) and v.get("_class") == "DataFrame":
            fr = _decode_frame(v)
            if k.startswith("res_"):
                net.stored_res[k] = fr
            else:
                net.tables[k] = fr
    for k, cols in _EMPTY_COLS.items():
        if k not in net.tables:
            net.tables[k] = Frame([], {c: np.empty(0, dtype=object) for c in cols})
    net.sn_mva = float(obj.get("sn_mva", 1.0))
    net.f_hz = float(obj.get("f_hz", 50.0))
    net.version = str(obj.get("version", ""))
    net.converged = bool(obj.get("converged", False))
    return net


# --------------------------------------------------------------------------------------
# pd2ppc: element tables -> per-unit bus/branch model  (pandapower/build_branch.py,
# build_bus.py, build_gen.py as published; MATPOWER case format)
# --------------------------------------------------------------------------------------
class PPC:
    pass


def _bus_lookup(net: Net):
    lab = net.bus.index
    look = {int(l): i for i, l in enumerate(lab)}
    return lab, look


def _line_params(net: Net, vn_of_label):
    """pandapower build_branch._calc_line_parameter: pi-model, per unit on the from-bus base."""
    ln = net.line
    n = len(ln)
    fb = ln.num("from_bus", 0).astype(np.int64)
    tb = ln.num("to_bus", 0).astype(np.int64)
    length = ln.num("length_km", 1.0)
    par = ln.num("parallel", 1.0)
    vn = np.array([vn_of_label(b) for b in fb], dtype=np.float64) if n else np.zeros(0)
    base_r = vn ** 2 / net.sn_mva
    r = ln.num("r_ohm_per_km", 0.0) * length / par / base_r
    x = ln.num("x_ohm_per_km", 0.0) * length / par / base_r
    b = 2.0 * np.pi * net.f_hz * ln.num("c_nf_per_km", 0.0) * 1e-9 * length * par * base_r
    g = ln.num("g_us_per_km", 0.0) * 1e-6 * length * par * base_r
    return fb, tb, r, x, g, b, np.ones(n), np.zeros(n), ln.flag("in_service")


def _trafo_params(net: Net, vn_of_label):
    """pandapower build_branch._calc_branch_values_from_trafo_df with trafo_model="t":
    tap changer on the hv/lv side (ratio only; every bundled grid has tap_phase_shifter False and
    tap_step_degree 0/NaN), short-circuit impedance referred to the lv bus base, magnetising branch
    from pfe_kw / i0_percent, T -> pi (wye-delta) conversion."""
    tr = net.trafo
    n = len(tr)
    hb = tr.num("hv_bus", 0).astype(np.int64)
    lb = tr.num("lv_bus", 0).astype(np.int64)
    if n == 0:
        z = np.zeros(0)
        return hb, lb, z, z, z, z, z, z, np.zeros(0, dtype=bool)
    vn_hv_bus = np.array([vn_of_label(b) for b in hb])
    vn_lv_bus = np.array([vn_of_label(b) for b in lb])
    vn_trafo_hv = tr.num("vn_hv_kv", 1.0).copy()
    vn_trafo_lv = tr.num("vn_lv_kv", 1.0).copy()
    tap_pos = tr.num("tap_pos", np.nan)
    tap_neutral = tr.num("tap_neutral", np.nan)
    tap_step = tr.num("tap_step_percent", np.nan)
    side = tr.text("tap_side")
    tap_diff = tap_pos - tap_neutral
    for i in range(n):
        if not np.isfinite(tap_diff[i]) or not np.isfinite(tap_step[i]):
            continue
        fac = 1.0 + tap_diff[i] * tap_step[i] / 100.0
        if side[i] == "hv":
            vn_trafo_hv[i] *= fac
        elif side[i] == "lv":
            vn_trafo_lv[i] *= fac
    par = tr.num("parallel", 1.0)
    sn_trafo = tr.num("sn_mva", 1.0)
    # _calc_r_x_from_dataframe
    tap_lv = (vn_trafo_lv / vn_lv_bus) ** 2 * net.sn_mva
    z_sc = tr.num("vk_percent", 0.0) / 100.0 / sn_trafo * tap_lv
    r_sc = tr.num("vkr_percent", 0.0) / 100.0 / sn_trafo * tap_lv
    x_sc = np.sign(z_sc) * np.sqrt(np.maximum(z_sc ** 2 - r_sc ** 2, 0.0))
    r = r_sc / par
    x = x_sc / par
    # _calc_y_from_dataframe: magnetising admittance  y_m = g_m + j b_m  (per unit, lv bus base)
    base_r = vn_lv_bus ** 2 / net.sn_mva
    vn_lv_kv = tr.num("vn_lv_kv", 1.0)
    pfe = tr.num("pfe_kw", 0.0) * 1e-3
    i0 = tr.num("i0_percent", 0.0)
    g_m = pfe / vn_lv_kv ** 2 * base_r
    b_sq = (i0 / 100.0 * sn_trafo) ** 2 - pfe ** 2
    b_sq[b_sq < 0] = 0.0
    b_m = -np.sign(i0) * np.sqrt(b_sq) * base_r / vn_lv_kv ** 2
    corr = (vn_trafo_lv / vn_lv_kv) ** 2
    g_m = g_m / corr * par
    b_m = b_m / corr * par
    # T -> pi
    g = np.zeros(n)
    b = np.zeros(n)
    for i in range(n):
        if g_m[i] == 0.0 and b_m[i] == 0.0:
            continue
        za = 0.5 * (r[i] + 1j * x[i])
        zc = 1.0 / (g_m[i] + 1j * b_m[i])
        zsum = za * za + 2.0 * za * zc
        zab = zsum / zc
        zbc = zsum / za
        r[i] = zab.real
        x[i] = zab.imag
        ysh = 2.0 / zbc           # total shunt admittance of the pi (half at each end)
        g[i] = ysh.real
        b[i] = ysh.imag
    ratio = (vn_trafo_hv / vn_trafo_lv) / (vn_hv_bus / vn_lv_bus)
    shift = tr.num("shift_degree", 0.0)
    return hb, lb, r, x, g, b, ratio, shift, tr.flag("in_service")


def _connected_to_ref(nb, f, t, stat, ref_idx):
    """Buses reachable from any reference bus over in-service branches (pandapower
    auxiliary._check_connectivity; used by rundcpp(check_connectivity=True))."""
    adj = [[] for _ in range(nb)]
    for a, b_, s in zip(f, t, stat):
        if s:
            adj[a].append(b_)
            adj[b_].append(a)
    seen = np.zeros(nb, dtype=bool)
    stack = [int(r) for r in ref_idx]
    for r in stack:
        seen[r] = True
    while stack:
        u = stack.pop()
        for v in adj[u]:
            if not seen[v]:
                seen[v] = True
                stack.append(v)
    return seen


def pd2ppc(net: Net) -> PPC:
    lab, look = _bus_lookup(net)
    nb = len(lab)
    vn = net.bus.num("vn_kv", 1.0)
    bus_is = net.bus.flag("in_service")

    def vn_of(label):
        return vn[look[int(label)]]

    p = PPC()
    p.base_mva = net.sn_mva
    p.nb = nb
    p.vn = vn
    p.bus_is = bus_is.copy()
    # ---- branches: lines then trafos (ppc branch order, pandapower build_branch._build_branch_ppc)
    lf, lt, lr, lx, lg, lbb, lratio, lshift, lstat = _line_params(net, vn_of)
    tf, tt, tr_, tx, tg, tb, tratio, tshift, tstat = _trafo_params(net, vn_of)
    p.n_line = len(lf)
    p.n_trafo = len(tf)
    f = np.array([look[int(b)] for b in np.concatenate([lf, tf])], dtype=np.int64)
    t = np.array([look[int(b)] for b in np.concatenate([lt, tt])], dtype=np.int64)
    p.f, p.t = f, t
    p.r = np.concatenate([lr, tr_])
    p.x = np.concatenate([lx, tx])
    p.g = np.concatenate([lg, tg])
    p.b = np.concatenate([lbb, tb])
    p.ratio = np.concatenate([lratio, tratio])
    p.shift = np.concatenate([lshift, tshift])
    stat = np.concatenate([lstat, tstat]).astype(bool)
    # a branch touching an out-of-service bus is out of service (pd2ppc._branches_with_oos_buses)
    stat &= bus_is[f] & bus_is[t] if len(f) else stat
    p.stat = stat
    # ---- bus injections: loads / storages / sgens (build_bus._calc_pq_elements_and_add_on_ppc)
    pd_ = np.zeros(nb)
    qd_ = np.zeros(nb)
    for name, sign in (("load", 1.0), ("storage", 1.0), ("sgen", -1.0)):
        tbl = net.tables[name]
        if len(tbl) == 0:
            continue
        b_ = np.array([look[int(v)] for v in tbl.num("bus", 0)], dtype=np.int64)
        ok = tbl.flag("in_service") & bus_is[b_]
        sc = tbl.num("scaling", 1.0)
        np.add.at(pd_, b_[ok], sign * (tbl.num("p_mw", 0.0) * sc)[ok])
        np.add.at(qd_, b_[ok], sign * (tbl.num("q_mvar", 0.0) * sc)[ok])
    p.pd, p.qd = pd_, qd_
    # ---- shunts (build_bus._calc_shunts_and_add_on_ppc): q>0 absorbs reactive power
    gs = np.zeros(nb)
    bs = np.zeros(nb)
    sh = net.shunt
    if len(sh):
        b_ = np.array([look[int(v)] for v in sh.num("bus", 0)], dtype=np.int64)
        ok = sh.flag("in_service") & bus_is[b_]
        vr = (vn[b_] / sh.num("vn_kv", 1.0)) ** 2
        step = sh.num("step", 1.0)
        np.add.at(gs, b_[ok], (sh.num("p_mw", 0.0) * step * vr)[ok])
        np.add.at(bs, b_[ok], (-sh.num("q_mvar", 0.0) * step * vr)[ok])
        p.sh_bus, p.sh_ok, p.sh_vr, p.sh_step = b_, ok, vr, step
    p.gs, p.bs = gs, bs
    # ---- generators: ext_grids first, then gens (build_gen._build_gen_ppc)
    eg, ge = net.ext_grid, net.gen
    eg_bus = np.array([look[int(v)] for v in eg.num("bus", 0)], dtype=np.int64)
    ge_bus = np.array([look[int(v)] for v in ge.num("bus", 0)], dtype=np.int64)
    eg_ok = eg.flag("in_service") & (bus_is[eg_bus] if len(eg_bus) else True)
    ge_ok = ge.flag("in_service") & (bus_is[ge_bus] if len(ge_bus) else True)
    p.n_eg, p.n_gen = len(eg), len(ge)
    p.gen_bus = np.concatenate([eg_bus, ge_bus]).astype(np.int64)
    p.gen_on = np.concatenate([eg_ok, ge_ok]).astype(bool)
    p.gen_pg = np.concatenate([np.zeros(len(eg)), ge.num("p_mw", 0.0) * ge.num("scaling", 1.0)])
    p.gen_vg = np.concatenate([eg.num("vm_pu", 1.0), ge.num("vm_pu", 1.0)])
    p.gen_qmin = np.concatenate([eg.num("min_q_mvar", -1e9), ge.num("min_q_mvar", -1e9)])
    p.gen_qmax = np.concatenate([eg.num("max_q_mvar", 1e9), ge.num("max_q_mvar", 1e9)])
    p.gen_is_ref = np.concatenate([np.ones(len(eg), dtype=bool), ge.flag("slack", False)])
    # bus types + initial voltages (flat magnitudes, generator set points on their buses)
    vm0 = np.ones(nb)
    va0 = np.zeros(nb)
    btype = np.full(nb, 1)                     # 1 = PQ, 2 = PV, 3 = REF, 4 = NONE
    btype[~bus_is] = 4
    for k in range(len(p.gen_bus)):            # table order: later generators win (fancy assignment)
        if p.gen_on[k]:
            vm0[p.gen_bus[k]] = p.gen_vg[k]
            if btype[p.gen_bus[k]] != 3:
                btype[p.gen_bus[k]] = 3 if p.gen_is_ref[k] else 2
    for k in range(len(eg)):
        if eg_ok[k]:
            va0[eg_bus[k]] = eg.num("va_degree", 0.0)[k]
    p.vm0, p.va0, p.btype = vm0, va0, btype
    return p


# --------------------------------------------------------------------------------------
# network matrices (PYPOWER makeYbus / makeBdc as used by pandapower)
# --------------------------------------------------------------------------------------
def make_ybus(p: PPC):
    nb = p.nb
    stat = p.stat.astype(float)
    with np.errstate(divide="ignore", invalid="ignore"):
        ys = stat / (p.r + 1j * p.x)
    ys[~p.stat] = 0.0
    bc = stat * (p.g + 1j * p.b)               # total line-charging admittance
    tap = p.ratio * np.exp(1j * np.pi / 180.0 * p.shift)
    ytt = ys + bc / 2.0
    yff = ytt / (tap * np.conj(tap))
    yft = -ys / np.conj(tap)
    ytf = -ys / tap
    Y = np.zeros((nb, nb), dtype=np.complex128)
    np.add.at(Y, (p.f, p.f), yff)
    np.add.at(Y, (p.f, p.t), yft)
    np.add.at(Y, (p.t, p.f), ytf)
    np.add.at(Y, (p.t, p.t), ytt)
    Y[np.arange(nb), np.arange(nb)] += (p.gs + 1j * p.bs) / p.base_mva
    return Y, (yff, yft, ytf, ytt)


def make_bdc(p: PPC):
    nb = p.nb
    with np.errstate(divide="ignore", invalid="ignore"):
        b = p.stat.astype(float) / p.x
    b[~p.stat] = 0.0
    b = b / p.ratio
    B = np.zeros((nb, nb))
    np.add.at(B, (p.f, p.f), b)
    np.add.at(B, (p.f, p.t), -b)
    np.add.at(B, (p.t, p.f), -b)
    np.add.at(B, (p.t, p.t), b)
    pfinj = b * (-p.shift * np.pi / 180.0)
    pbusinj = np.zeros(nb)
    np.add.at(pbusinj, p.f, pfinj)
    np.add.at(pbusinj, p.t, -pfinj)
    return B, b, pfinj, pbusinj


def make_sbus(p: PPC):
    s = -(p.pd + 1j * p.qd)
    on = p.gen_on
    sg = np.zeros(p.nb, dtype=np.complex128)
    np.add.at(sg, p.gen_bus[on], p.gen_pg[on].astype(np.complex128))
    return (s + sg) / p.base_mva


def _dc_angles(p: PPC, live, ref, va_start):
    """PYPOWER dcpf: B[pvpq,pvpq] Va[pvpq] = Pbus[pvpq] - B[pvpq,ref] Va[ref]."""
    B, b, pfinj, pbusinj = make_bdc(p)
    pbus = make_sbus(p).real - pbusinj - p.gs / p.base_mva
    va = va_start.copy()
    rest = np.array([i for i in live if i not in set(ref.tolist())], dtype=np.int64)
    if len(rest):
        rhs = pbus[rest] - B[np.ix_(rest, ref)] @ va[ref]
        try:
            va[rest] = np.linalg.solve(B[np.ix_(rest, rest)], rhs)
        except np.linalg.LinAlgError:
            va[rest] = np.nan
    return va, b, pfinj


def _split_q(p: PPC, on_idx, qg_bus_total):
    """PYPOWER pfsoln: a bus's reactive dispatch is shared between its units in proportion to
    their reactive range (equal shares when the range is degenerate)."""
    qg = np.zeros(len(p.gen_bus))
    nb = p.nb
    cnt = np.zeros(nb)
    np.add.at(cnt, p.gen_bus[on_idx], 1.0)
    qmin_b = np.zeros(nb)
    qmax_b = np.zeros(nb)
    np.add.at(qmin_b, p.gen_bus[on_idx], p.gen_qmin[on_idx])
    np.add.at(qmax_b, p.gen_bus[on_idx], p.gen_qmax[on_idx])
    eps = np.finfo(float).eps
    for k in on_idx:
        bb = p.gen_bus[k]
        tot = qg_bus_total[bb]
        if cnt[bb] <= 1 or qmin_b[bb] == qmax_b[bb]:
            qg[k] = tot / cnt[bb]
        else:
            qg[k] = p.gen_qmin[k] + (tot - qmin_b[bb]) / (qmax_b[bb] - qmin_b[bb] + eps) * (p.gen_qmax[k] - p.gen_qmin[k])
    return qg


def _slack_p(p: PPC, on_idx, ref, pinj_bus):
    """pandapower pfsoln: the slack power of a reference bus is shared equally between the
    reference units (ext_grids + slack generators) sitting on it; other units keep their set point."""
    pg = p.gen_pg.copy()
    pg[~p.gen_on] = 0.0
    for rb in ref:
        at_bus = [k for k in on_idx if p.gen_bus[k] == rb]
        refs = [k for k in at_bus if p.gen_is_ref[k]]
        others = [k for k in at_bus if not p.gen_is_ref[k]]
        if not refs:
            continue
        total = pinj_bus[rb] * p.base_mva + p.pd[rb]
        share = (total - sum(pg[k] for k in others)) / len(refs)
        for k in refs:
            pg[k] = share
    return pg


# --------------------------------------------------------------------------------------
# solvers
# --------------------------------------------------------------------------------------
def _types(p: PPC, live_mask):
    ref = np.array([i for i in range(p.nb) if live_mask[i] and p.btype[i] == 3], dtype=np.int64)
    pv = np.array([i for i in range(p.nb) if live_mask[i] and p.btype[i] == 2], dtype=np.int64)
    pq = np.array([i for i in range(p.nb) if live_mask[i] and p.btype[i] == 1], dtype=np.int64)
    return ref, pv, pq


def newton_raphson(Y, sbus, v0, ref, pv, pq, tol, max_it):
    """PYPOWER / pandapower ``newtonpf`` in polar coordinates (full Newton, dense LU here)."""
    v = v0.copy()
    va = np.angle(v)
    vm = np.abs(v)
    pvpq = np.concatenate([pv, pq]).astype(np.int64)
    npv, npq = len(pv), len(pq)

    def mism(vv):
        mis = vv * np.conj(Y @ vv) - sbus
        return np.concatenate([mis[pvpq].real, mis[pq].imag])

    F = mism(v)
    it = 0
    conv = bool(np.all(np.isfinite(F))) and (len(F) == 0 or np.max(np.abs(F)) < tol)
    while not conv and it < max_it:
        it += 1
        ibus = Y @ v
        diag_v = np.diag(v)
        diag_i = np.diag(ibus)
        diag_vn = np.diag(v / np.abs(v))
        ds_dvm = diag_v @ np.conj(Y @ diag_vn) + np.conj(diag_i) @ diag_vn
        ds_dva = 1j * diag_v @ np.conj(diag_i - Y @ diag_v)
        J = np.block([
            [ds_dva[np.ix_(pvpq, pvpq)].real, ds_dvm[np.ix_(pvpq, pq)].real],
            [ds_dva[np.ix_(pq, pvpq)].imag, ds_dvm[np.ix_(pq, pq)].imag],
        ])
        try:
            dx = -np.linalg.solve(J, F)
        except np.linalg.LinAlgError:
            return v, False, it
        if not np.all(np.isfinite(dx)):
            return v, False, it
        va[pv] += dx[:npv]
        va[pq] += dx[npv:npv + npq]
        vm[pq] += dx[npv + npq:]
        v = vm * np.exp(1j * va)
        vm = np.abs(v)
        va = np.angle(v)
        F = mism(v)
        conv = bool(np.all(np.isfinite(F))) and np.max(np.abs(F)) < tol
    return v, conv, it


def _run(net: Net, ac: bool, check_connectivity: bool, max_iteration=10, tolerance_mva=1e-8):
    p = pd2ppc(net)
    live = p.bus_is.copy()
    ref_all = np.array([i for i in range(p.nb) if live[i] and p.btype[i] == 3], dtype=np.int64)
    if len(ref_all) == 0:
        raise LoadflowNotConverged("no reference bus available")
    reach = _connected_to_ref(p.nb, p.f, p.t, p.stat, ref_all)
    if check_connectivity:
        # unsupplied buses and everything on them are taken out of service; their results are NaN
        live &= reach
        p.stat &= live[p.f] & live[p.t]
        p.gen_on &= live[p.gen_bus]
        p.pd[~live] = 0.0
        p.qd[~live] = 0.0
        p.gs[~live] = 0.0
        p.bs[~live] = 0.0
    elif not np.all(reach[live]):
        # check_connectivity=False: the admittance / Jacobian matrix of a grid with an
        # unsupplied in-service bus is singular -> pandapower's Newton never converges
        raise LoadflowNotConverged("in-service bus not connected to a reference bus (singular system)")
    ref, pv, pq = _types(p, live)
    live_idx = np.flatnonzero(live)
    va_dc, bdc, pfinj = _dc_angles(p, live_idx, ref, np.deg2rad(p.va0))
    Y, (yff, yft, ytf, ytt) = make_ybus(p)
    on_idx = np.flatnonzero(p.gen_on)
    res = PPC()
    res.p = p
    res.live = live
    if not ac:
        if not np.all(np.isfinite(va_dc[live_idx])):
            raise LoadflowNotConverged("DC system singular")
        pf = (bdc * (va_dc[p.f] - va_dc[p.t]) + pfinj) * p.base_mva
        pf[~p.stat] = 0.0
        res.sf = pf.astype(np.complex128)
        res.st = -pf.astype(np.complex128)
        res.vm = p.vm0.copy()
        res.va = va_dc
        B, _, _, pbusinj = make_bdc(p)
        pinj = B @ np.where(live, va_dc, 0.0) + pbusinj + p.gs / p.base_mva
        res.pg = _slack_p(p, on_idx, ref, pinj)
        res.qg = np.zeros(len(p.gen_bus))
        res.iters = 0
        res.ac = False
        return res
    if not np.all(np.isfinite(va_dc[live_idx])):
        raise LoadflowNotConverged("DC initialisation failed (singular)")
    v0 = p.vm0 * np.exp(1j * va_dc)
    v0[~live] = 0.0
    sbus = make_sbus(p)
    idx = live_idx
    sub = {int(g): i for i, g in enumerate(idx)}
    Yl = Y[np.ix_(idx, idx)]
    v, conv, it = newton_raphson(Yl, sbus[idx], v0[idx],
                                 np.array([sub[int(i)] for i in ref], dtype=np.int64),
                                 np.array([sub[int(i)] for i in pv], dtype=np.int64),
                                 np.array([sub[int(i)] for i in pq], dtype=np.int64),
                                 tolerance_mva / p.base_mva, max_iteration)
    if not conv:
        raise LoadflowNotConverged(f"Power Flow nr did not converge after {max_iteration} iterations!")
    V = np.zeros(p.nb, dtype=np.complex128)
    V[idx] = v
    sf = V[p.f] * np.conj(yff * V[p.f] + yft * V[p.t]) * p.base_mva
    st = V[p.t] * np.conj(ytf * V[p.f] + ytt * V[p.t]) * p.base_mva
    sf[~p.stat] = 0.0
    st[~p.stat] = 0.0
    sinj = V * np.conj(Y @ V)
    res.sf, res.st = sf, st
    res.vm = np.abs(V)
    res.va = np.angle(V)
    res.pg = _slack_p(p, on_idx, ref, sinj.real)
    qtot = sinj.imag * p.base_mva + p.qd
    res.qg = _split_q(p, on_idx, qtot)
    res.iters = it
    res.ac = True
    return res


def _fill_results(net: Net, r):
    """The ``res_*`` tables PandaPowerBackend reads (pandaPowerBackend.py:1122-1218, 1526-1647).
    Column semantics follow pandapower/results_branch.py, results_bus.py, results_gen.py."""
    p = r.p
    lab, look = _bus_lookup(net)
    nan = np.nan
    vm_b = np.where(r.live, r.vm, nan)
    va_b = np.where(r.live, np.rad2deg(r.va), nan)
    res = {}
    res["bus"] = {"vm_pu": (vm_b.copy() if r.ac else np.full(p.nb, nan)), "va_degree": va_b.copy()}
    nl, nt = p.n_line, p.n_trafo
    vkv = r.vm * p.vn

    def branch(sl, names):
        f, t = p.f[sl], p.t[sl]
        sf, st = r.sf[sl], r.st[sl]
        mag_f = np.abs(sf) if r.ac else np.abs(sf.real)
        mag_t = np.abs(st) if r.ac else np.abs(st.real)
        with np.errstate(divide="ignore", invalid="ignore"):
            i_f = mag_f / (SQRT3 * vkv[f])
            i_t = mag_t / (SQRT3 * vkv[t])
        out = {
            names[0]: sf.real.copy(), names[1]: (sf.imag.copy() if r.ac else np.zeros(len(f))),
            names[2]: st.real.copy(), names[3]: (st.imag.copy() if r.ac else np.zeros(len(f))),
            names[4]: i_f, names[5]: i_t,
            names[6]: np.where(r.live[f], r.vm[f], nan), names[7]: np.where(r.live[f], np.rad2deg(r.va[f]), nan),
            names[8]: np.where(r.live[t], r.vm[t], nan), names[9]: np.where(r.live[t], np.rad2deg(r.va[t]), nan),
        }
        return out

    res["line"] = branch(slice(0, nl), ["p_from_mw", "q_from_mvar", "p_to_mw", "q_to_mvar", "i_from_ka", "i_to_ka",
                                        "vm_from_pu", "va_from_degree", "vm_to_pu", "va_to_degree"])
    res["trafo"] = branch(slice(nl, nl + nt), ["p_hv_mw", "q_hv_mvar", "p_lv_mw", "q_lv_mvar", "i_hv_ka", "i_lv_ka",
                                               "vm_hv_pu", "va_hv_degree", "vm_lv_pu", "va_lv_degree"])
    ne, ng = p.n_eg, p.n_gen
    pg = np.where(p.gen_on, r.pg, 0.0)
    qg = np.where(p.gen_on, r.qg, 0.0)
    gb = p.gen_bus
    res["ext_grid"] = {"p_mw": pg[:ne].copy(), "q_mvar": qg[:ne].copy()}
    res["gen"] = {"p_mw": pg[ne:].copy(), "q_mvar": qg[ne:].copy(),
                  # results_gen._get_v_gen_results: zeros for units that are not in service
                  "vm_pu": np.where(p.gen_on[ne:], r.vm[gb[ne:]], 0.0),
                  "va_degree": np.where(p.gen_on[ne:], np.rad2deg(r.va[gb[ne:]]), 0.0)}
    net.ppc_internal = {"gen_pg": r.pg.copy(), "gen_qg": r.qg.copy()}
    for name in ("load", "storage"):
        tbl = net.tables[name]
        if len(tbl):
            b_ = np.array([look[int(v)] for v in tbl.num("bus", 0)], dtype=np.int64)
            ok = tbl.flag("in_service") & r.live[b_]
            sc = tbl.num("scaling", 1.0)
            res[name] = {"p_mw": np.where(ok, tbl.num("p_mw", 0.0) * sc, 0.0),
                         "q_mvar": np.where(ok, tbl.num("q_mvar", 0.0) * sc, 0.0) if r.ac else np.zeros(len(tbl))}
        else:
            res[name] = {"p_mw": np.zeros(0), "q_mvar": np.zeros(0)}
    sh = net.shunt
    if len(sh):
        b_ = p.sh_bus
        ok = sh.flag("in_service") & r.live[b_]
        v2 = np.where(ok, r.vm[b_] ** 2, 0.0) if r.ac else np.where(ok, 1.0, 0.0)
        res["shunt"] = {"p_mw": sh.num("p_mw", 0.0) * p.sh_step * p.sh_vr * v2,
                        "q_mvar": (sh.num("q_mvar", 0.0) * p.sh_step * p.sh_vr * v2) if r.ac else np.zeros(len(sh)),
                        "vm_pu": np.where(r.live[b_], r.vm[b_], nan)}
    else:
        res["shunt"] = {"p_mw": np.zeros(0), "q_mvar": np.zeros(0), "vm_pu": np.zeros(0)}
    net.res = res
    net.converged = True
    net.iterations = r.iters


def runpp(net: Net, max_iteration: int = 10, tolerance_mva: float = 1e-8, check_connectivity: bool = False):
    """``pp.runpp(net, init="dc", check_connectivity=False, max_iteration=10, distributed_slack=False)``."""
    net.converged = False
    r = _run(net, True, check_connectivity, max_iteration, tolerance_mva)
    _fill_results(net, r)


def rundcpp(net: Net, check_connectivity: bool = True):
    """``pp.rundcpp(net, check_connectivity=True, init="flat")``."""
    net.converged = False
    r = _run(net, False, check_connectivity)
    _fill_results(net, r)
