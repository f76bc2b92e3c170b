"""ORACLE (test infrastructure, never shipped / never on the product path).

Restatement of ``grid2op.Backend.PandaPowerBackend`` (reference:
grid2op/Backend/pandaPowerBackend.py, cited per method below) on top of the pandapower
restatement in :mod:`oracle.pandapower_ref`.  It is a real ``grid2op.Backend.Backend`` subclass so
the reference's own backend test-suites (``grid2op._create_test_suite.create_test_suite``) and
golden vectors (grid2op/tests/BaseBackendTest.py:258-319, 438-530;
grid2op/data/rte_case5_example/_statistics) can be run against it on the CPU.

The product backend (``grid2op_b200.B200Backend``) shares no code with this file; tests compare
the two on identical action sequences.
"""
from __future__ import annotations

import copy
import os
from typing import Optional, Tuple, Union

import numpy as np

from grid2op_b200._bootstrap import ensure_grid2op

if not ensure_grid2op():  # pragma: no cover
    raise ImportError("grid2op is required by the oracle backend")

from grid2op.Backend.backend import Backend  # noqa: E402
from grid2op.dtypes import dt_bool, dt_float, dt_int  # noqa: E402
from grid2op.Exceptions import BackendError  # noqa: E402

from . import pandapower_ref as ppr  # noqa: E402


class PandaPowerBackendRef(Backend):
    shunts_data_available = True

    # pandaPowerBackend.py:119-252
    def __init__(self, detailed_infos_for_cascading_failures: bool = False, can_be_copied: bool = True,
                 max_iter: int = 10):
        Backend.__init__(self, detailed_infos_for_cascading_failures=detailed_infos_for_cascading_failures,
                         can_be_copied=can_be_copied, max_iter=max_iter)
        self._needs_active_bus = True
        self._max_iter = max_iter
        self.can_output_theta = True
        self._net: Optional[ppr.Net] = None
        self._net0: Optional[ppr.Net] = None
        self._iref_slack = None
        self._id_bus_added = None
        self._n_true_line = -1
        self._topo_vect = None
        self.div_exception = None
        self.tol = 1e-5

    # ------------------------------------------------------------------ load_grid (pPB:356-617)
    def _pf_init_run(self):
        try:
            ppr.runpp(self._net, max_iteration=self._max_iter)
        except ppr.LoadflowNotConverged:
            ppr.rundcpp(self._net)

    def load_grid(self, path, filename=None) -> None:
        self.can_handle_more_than_2_busbar()
        self.can_handle_detachment()
        full = self.make_complete_path(path, filename)
        net = ppr.from_json(full)
        self._net = net
        self._pf_init_run()
        self._iref_slack = None
        self._id_bus_added = None
        slack_flags = net.gen.flag("slack", False)
        if len(net.gen) == 0 or not slack_flags.any():
            # pPB:394-451: a generator is created on every reference unit's bus that carries no generator
            ppcb = ppr.pd2ppc(net)
            gen_buses = set(int(b) for b in net.gen.num("bus", -1))
            lab = net.bus.index
            i_ref = None
            for k in range(ppcb.n_eg + ppcb.n_gen):
                bus_lab = int(lab[ppcb.gen_bus[k]])
                if bus_lab in gen_buses:
                    continue
                new_id = len(net.gen)
                net.gen.append_row(
                    (int(net.gen.index.max()) + 1) if len(net.gen) else 0,
                    bus=bus_lab, p_mw=float(net.ppc_internal["gen_pg"][k]), vm_pu=float(ppcb.gen_vg[k]),
                    min_q_mvar=float(ppcb.gen_qmin[k]), max_q_mvar=float(ppcb.gen_qmax[k]),
                    scaling=1.0, slack=(i_ref is None), in_service=True, slack_weight=1.0, controllable=True)
                if i_ref is None:
                    i_ref = k
                    self._iref_slack = k
                    self._id_bus_added = new_id
                    eg = net.ext_grid
                    net.tables["ext_grid"] = ppr.Frame(eg.index[:1], {c: v[:1] for c, v in eg.cols.items()})
        self._pf_init_run()
        n_sub = len(net.bus)
        self._n_true_line = len(net.line)
        self._init_bus_load = net.load.num("bus", 0).astype(np.int64)
        self._init_bus_gen = net.gen.num("bus", 0).astype(np.int64)
        self._init_bus_lor = np.concatenate([net.line.num("from_bus", 0), net.trafo.num("hv_bus", 0)]).astype(np.int64)
        self._init_bus_lex = np.concatenate([net.line.num("to_bus", 0), net.trafo.num("lv_bus", 0)]).astype(np.int64)
        if len(net.ext_grid):
            net.ext_grid["va_degree"] = np.array([0.0] * len(net.ext_grid), dtype=object)

        # names (pPB:481-553)
        self.n_line = len(net.line) + len(net.trafo)
        ln_names = net.line.text("name")
        if "name" in net.line and all(v is not None for v in ln_names):
            names = list(ln_names)
        else:
            names = [f"{int(f)}_{int(t)}_{i}" for i, (f, t) in
                     enumerate(zip(net.line.num("from_bus", 0), net.line.num("to_bus", 0)))]
        tr_names = net.trafo.text("name")
        if "name" in net.trafo and len(net.trafo) and all(v is not None for v in tr_names):
            names += list(tr_names)
        else:
            for i, (h, l) in enumerate(zip(net.trafo.num("hv_bus", 0), net.trafo.num("lv_bus", 0))):
                a, b = sorted((str(int(h)), str(int(l))))
                names.append(f"{a}_{b}_{i + len(net.line)}")
        self.name_line = np.array(names)

        def _names(tbl, prefix):
            raw = tbl.text("name")
            if "name" in tbl and len(tbl) and all(v is not None for v in raw):
                return np.array(list(raw))
            return np.array([f"{prefix}_{int(b)}_{i}" for i, b in enumerate(tbl.num("bus", 0))])

        self.n_gen = len(net.gen)
        self.name_gen = _names(net.gen, "gen")
        self.n_load = len(net.load)
        self.name_load = _names(net.load, "load")
        self.n_storage = len(net.storage)
        if self.n_storage == 0:
            self.set_no_storage()
        else:
            self.name_storage = _names(net.storage, "storage")
        self.n_sub = n_sub
        self.name_sub = np.array([f"sub_{int(i)}" for i in net.bus.index])
        self.n_shunt = len(net.shunt)

        # extra busbars (pPB:561-576): copies of every bus, label += n_sub, out of service
        base = net.bus.copy()
        for k in range(1, self.n_busbar_per_sub):
            for i in range(n_sub):
                vals = {c: base.cols[c][i] for c in base.cols}
                vals["in_service"] = False
                net.bus.append_row(int(base.index[i]) + k * n_sub, **vals)
        self._pf_init_run()
        self._init_private_attrs()
        self.comp_time = 0.0

    # pPB:670-874
    def _init_private_attrs(self):
        net = self._net
        self.sub_info = np.zeros(self.n_sub, dtype=dt_int)
        self.load_to_subid = np.zeros(self.n_load, dtype=dt_int)
        self.gen_to_subid = np.zeros(self.n_gen, dtype=dt_int)
        self.line_or_to_subid = np.zeros(self.n_line, dtype=dt_int)
        self.line_ex_to_subid = np.zeros(self.n_line, dtype=dt_int)
        self.load_to_sub_pos = np.zeros(self.n_load, dtype=dt_int)
        self.gen_to_sub_pos = np.zeros(self.n_gen, dtype=dt_int)
        self.line_or_to_sub_pos = np.zeros(self.n_line, dtype=dt_int)
        self.line_ex_to_sub_pos = np.zeros(self.n_line, dtype=dt_int)
        if self.n_storage > 0:
            self.storage_to_subid = np.zeros(self.n_storage, dtype=dt_int)
            self.storage_to_sub_pos = np.zeros(self.n_storage, dtype=dt_int)
        used = np.zeros(self.n_sub, dtype=dt_int)

        def place(sub):
            pos = used[sub]
            used[sub] += 1
            self.sub_info[sub] += 1
            return pos

        for i in range(self.n_line):
            so, se = int(self._init_bus_lor[i]), int(self._init_bus_lex[i])
            self.line_or_to_subid[i] = so
            self.line_ex_to_subid[i] = se
            self.line_or_to_sub_pos[i] = place(so)
            self.line_ex_to_sub_pos[i] = place(se)
        for i, b in enumerate(self._init_bus_gen):
            self.gen_to_subid[i] = b
            self.gen_to_sub_pos[i] = place(int(b))
        for i, b in enumerate(self._init_bus_load):
            self.load_to_subid[i] = b
            self.load_to_sub_pos[i] = place(int(b))
        if self.n_storage > 0:
            for i, b in enumerate(net.storage.num("bus", 0).astype(int)):
                self.storage_to_subid[i] = b
                self.storage_to_sub_pos[i] = place(int(b))
        self.dim_topo = int(self.sub_info.sum())
        sh_bus = net.shunt.num("bus", 0).astype(int)
        self.shunt_to_subid = np.array(sh_bus, dtype=dt_int)
        self.name_shunt = np.array([f"shunt_{int(b)}_{i}" for i, b in enumerate(sh_bus)]).astype(str)
        vn = net.bus.num("vn_kv", 1.0)
        look = {int(l): i for i, l in enumerate(net.bus.index)}

        def vn_of(subs):
            return np.array([vn[look[int(s)]] for s in subs], dtype=np.float64)

        self._sh_vnkv = vn_of(self.shunt_to_subid)
        self._compute_pos_big_topo()
        self._topo_vect = np.full(self.dim_topo, -1, dtype=dt_int)
        self.load_pu_to_kv = vn_of(self.load_to_subid).astype(dt_float)
        self.prod_pu_to_kv = vn_of(self.gen_to_subid).astype(dt_float)
        self.lines_or_pu_to_kv = vn_of(self.line_or_to_subid).astype(dt_float)
        self.lines_ex_pu_to_kv = vn_of(self.line_ex_to_subid).astype(dt_float)
        self.storage_pu_to_kv = vn_of(self.storage_to_subid).astype(dt_float) if self.n_storage else np.zeros(0, dt_float)
        self.thermal_limit_a = (1000.0 * np.concatenate([
            net.line.num("max_i_ka", 0.0),
            net.trafo.num("sn_mva", 0.0) / (np.sqrt(3) * net.trafo.num("vn_hv_kv", 1.0))])).astype(dt_float)
        for nm, n in (("p_or", self.n_line), ("q_or", self.n_line), ("v_or", self.n_line), ("a_or", self.n_line),
                      ("p_ex", self.n_line), ("q_ex", self.n_line), ("v_ex", self.n_line), ("a_ex", self.n_line),
                      ("theta_or", self.n_line), ("theta_ex", self.n_line),
                      ("load_p", self.n_load), ("load_q", self.n_load), ("load_v", self.n_load), ("load_theta", self.n_load),
                      ("prod_p", self.n_gen), ("prod_q", self.n_gen), ("prod_v", self.n_gen), ("gen_theta", self.n_gen),
                      ("storage_p", self.n_storage), ("storage_q", self.n_storage), ("storage_v", self.n_storage),
                      ("storage_theta", self.n_storage)):
            setattr(self, nm, np.full(n, np.nan, dtype=dt_float))
        self.line_status = np.zeros(self.n_line, dtype=dt_bool)
        self._refresh_status_topo()
        self._net0 = self._net.deepcopy()

    # ------------------------------------------------------------------ helpers
    def _l2g(self, local, sub):
        return -1 if local == -1 else int(sub) + (int(local) - 1) * self.n_sub

    def _g2l(self, glob, ok):
        glob = np.asarray(glob, dtype=np.int64)
        res = glob // self.n_sub + 1
        res[~np.asarray(ok, dtype=bool)] = -1
        return res.astype(dt_int)

    # pPB:1450-1459 + 1489-1524
    def _refresh_status_topo(self):
        cls = type(self)
        net = self._net
        st = np.concatenate([net.line.flag("in_service"), net.trafo.flag("in_service")])
        self.line_status = st.astype(dt_bool)
        tv = np.full(self.dim_topo if self.dim_topo is None or self.dim_topo < 0 else self.dim_topo, -1, dtype=dt_int)
        lor_pos = self.line_or_pos_topo_vect if self.line_or_pos_topo_vect is not None else self.line_or_pos_topo_vect
        lex_pos = self.line_ex_pos_topo_vect if self.line_ex_pos_topo_vect is not None else self.line_ex_pos_topo_vect
        tv[lor_pos] = self._g2l(np.concatenate([net.line.num("from_bus", 0), net.trafo.num("hv_bus", 0)]), st)
        tv[lex_pos] = self._g2l(np.concatenate([net.line.num("to_bus", 0), net.trafo.num("lv_bus", 0)]), st)
        tv[self.load_pos_topo_vect] = self._g2l(net.load.num("bus", 0), net.load.flag("in_service"))
        tv[self.gen_pos_topo_vect] = self._g2l(net.gen.num("bus", 0), net.gen.flag("in_service"))
        if self.n_storage:
            tv[self.storage_pos_topo_vect] = self._g2l(net.storage.num("bus", 0), net.storage.flag("in_service"))
        self._topo_vect = tv

    # ------------------------------------------------------------------ apply_action (pPB:902-1067)
    def apply_action(self, backend_action) -> None:
        cls = type(self)
        net = self._net
        active_bus, (prod_p, prod_v, load_p, load_q, storage), topo__, shunts__ = backend_action()
        flat = np.asarray(active_bus).T.reshape(-1)
        # the reference assigns a Series indexed 0..n-1 -> aligned on the bus *labels* (pPB:920-922)
        net.bus["in_service"] = np.array([bool(flat[int(l)]) for l in net.bus.index], dtype=object)
        for i in np.flatnonzero(prod_p.changed):
            net.gen["p_mw"][i] = float(prod_p.values[i])
        for i in np.flatnonzero(prod_v.changed):
            # float32 / float32 division, as the reference does (pPB:927): the set point is rounded to float32
            net.gen["vm_pu"][i] = float(np.float32(prod_v.values[i]) / np.float32(self.prod_pu_to_kv[i]))
        if self._id_bus_added is not None and prod_v.changed[self._id_bus_added]:
            net.ext_grid["vm_pu"][:] = float(net.gen["vm_pu"][self._id_bus_added])
        for i in np.flatnonzero(load_p.changed):
            net.load["p_mw"][i] = float(load_p.values[i])
        for i in np.flatnonzero(load_q.changed):
            net.load["q_mvar"][i] = float(load_q.values[i])
        if self.n_storage > 0:
            for i in np.flatnonzero(storage.changed):
                net.storage["p_mw"][i] = float(storage.values[i])
            sb = backend_action.get_storages_bus()
            for i in np.flatnonzero(sb.changed):
                g = self._l2g(int(sb.values[i]), self.storage_to_subid[i])
                if g <= -1:
                    net.storage["in_service"][i] = False
                    net.storage["bus"][i] = int(self.storage_to_subid[i])
                else:
                    net.storage["in_service"][i] = True
                    net.storage["bus"][i] = g
        if self.shunts_data_available:
            sh_p, sh_q, sh_b = shunts__
            for i in np.flatnonzero(sh_p.changed):
                net.shunt["p_mw"][i] = float(sh_p.values[i])
            for i in np.flatnonzero(sh_q.changed):
                net.shunt["q_mvar"][i] = float(sh_q.values[i])
            for i in np.flatnonzero(sh_b.changed):
                if int(sh_b.values[i]) == -1:
                    net.shunt["in_service"][i] = False
                else:
                    net.shunt["in_service"][i] = True
                    net.shunt["bus"][i] = self._l2g(int(sh_b.values[i]), self.shunt_to_subid[i])
        nl = self._n_true_line
        kind = self._pos_kind
        for pos, new_bus in topo__:
            what, idx = kind[pos]
            new_bus = int(new_bus)
            if what == "load":
                self._set_bus(net.load, idx, "bus", self._l2g(new_bus, self._init_bus_load[idx]))
            elif what == "gen":
                g = self._l2g(new_bus, self._init_bus_gen[idx])
                self._set_bus(net.gen, idx, "bus", g)
                if g >= 0 and idx == len(net.gen) - 1 and self._iref_slack is not None:
                    net.ext_grid["bus"][0] = g           # pPB:997-1003
            elif what == "lor":
                g = self._l2g(new_bus, self._init_bus_lor[idx])
                if idx < nl:
                    self._set_bus(net.line, idx, "from_bus", g)
                else:
                    self._set_bus(net.trafo, idx - nl, "hv_bus", g)
            elif what == "lex":
                g = self._l2g(new_bus, self._init_bus_lex[idx])
                if idx < nl:
                    self._set_bus(net.line, idx, "to_bus", g)
                else:
                    self._set_bus(net.trafo, idx - nl, "lv_bus", g)

    @staticmethod
    def _set_bus(tbl, idx, col, glob):   # pPB:1057-1067
        if glob >= 0:
            tbl["in_service"][idx] = True
            tbl[col][idx] = glob
        else:
            tbl["in_service"][idx] = False

    @property
    def _pos_kind(self):
        if getattr(self, "_pos_kind_cache", None) is None:
            cls = type(self)
            kind = [(None, None)] * self.dim_topo
            for i, pos in enumerate(self.load_pos_topo_vect):
                kind[pos] = ("load", i)
            for i, pos in enumerate(self.gen_pos_topo_vect):
                kind[pos] = ("gen", i)
            for i, pos in enumerate(self.line_or_pos_topo_vect):
                kind[pos] = ("lor", i)
            for i, pos in enumerate(self.line_ex_pos_topo_vect):
                kind[pos] = ("lex", i)
            self._pos_kind_cache = kind
        return self._pos_kind_cache

    # ------------------------------------------------------------------ runpf (pPB:1078-1120, 1220-1255)
    def runpf(self, is_dc: bool = False) -> Tuple[bool, Union[Exception, None]]:
        net = self._net
        try:
            self._refresh_status_topo()
            if is_dc:
                ppr.rundcpp(net, check_connectivity=True)
            else:
                ppr.runpp(net, max_iteration=self._max_iter, check_connectivity=False)
            self.comp_time += 1e-6
            rg = net.res["gen"]
            if any(np.isnan(rg[c]).any() for c in ("p_mw", "q_mvar", "vm_pu", "va_degree")):
                raise ppr.LoadflowNotConverged("Divergence due to Nan values in res_gen table")
            bus_ok = net.bus.flag("in_service")
            if np.isnan(net.res["bus"]["va_degree"][bus_ok]).any():
                raise ppr.LoadflowNotConverged("Isolated bus")
            self._fetch(is_dc)
            return True, None
        except ppr.LoadflowNotConverged as exc_:
            self.div_exception = exc_
            self._all_nan()
            return False, BackendError(f'powerflow diverged with error :"{exc_}"')

    # pPB:1122-1218, 1526-1564, 1621-1647
    def _fetch(self, is_dc):
        cls = type(self)
        net = self._net
        res = net.res
        lab = {int(l): i for i, l in enumerate(net.bus.index)}
        bus_vm, bus_va = res["bus"]["vm_pu"], res["bus"]["va_degree"]
        prod_p = res["gen"]["p_mw"].astype(dt_float).copy()
        prod_q = res["gen"]["q_mvar"].astype(dt_float).copy()
        if self._iref_slack is not None:
            prod_p[self._id_bus_added] += net.ppc_internal["gen_pg"][self._iref_slack]
            prod_q[self._id_bus_added] += net.ppc_internal["gen_qg"][self._iref_slack]
        self.prod_p[:] = prod_p
        self.prod_q[:] = prod_q
        self.prod_v[:] = res["gen"]["vm_pu"].astype(dt_float) * self.prod_pu_to_kv
        self.gen_theta[:] = res["gen"]["va_degree"].astype(dt_float)
        lb = np.array([lab[int(b)] for b in net.load.num("bus", 0)], dtype=np.int64)
        l_on = net.load.flag("in_service")
        self.load_p[:] = res["load"]["p_mw"].astype(dt_float)
        self.load_q[:] = res["load"]["q_mvar"].astype(dt_float)
        lv = bus_vm[lb].astype(dt_float) * self.load_pu_to_kv
        lth = bus_va[lb].astype(dt_float)
        lv[~l_on] = 0.0
        lth[~l_on] = 0.0
        self.load_v[:] = lv
        self.load_theta[:] = lth
        if is_dc:
            self.load_v[:] = self.load_pu_to_kv
            for l_id in range(self.n_load):
                same = np.flatnonzero(self.gen_to_subid == self.load_to_subid[l_id])
                for g_id in same:
                    if self._topo_vect[self.load_pos_topo_vect[l_id]] == self._topo_vect[self.gen_pos_topo_vect[g_id]]:
                        self.load_v[l_id] = self.prod_v[g_id]
                        break
            self.load_v[~l_on] = 0.0

        def cat(a, b):
            return np.concatenate([res["line"][a], res["trafo"][b]])

        self.p_or[:] = cat("p_from_mw", "p_hv_mw")
        self.q_or[:] = cat("q_from_mvar", "q_hv_mvar")
        self.v_or[:] = cat("vm_from_pu", "vm_hv_pu")
        self.a_or[:] = cat("i_from_ka", "i_hv_ka") * 1000.0
        self.theta_or[:] = cat("va_from_degree", "va_hv_degree")
        self.p_ex[:] = cat("p_to_mw", "p_lv_mw")
        self.q_ex[:] = cat("q_to_mvar", "q_lv_mvar")
        self.v_ex[:] = cat("vm_to_pu", "vm_lv_pu")
        self.a_ex[:] = cat("i_to_ka", "i_lv_ka") * 1000.0
        self.theta_ex[:] = cat("va_to_degree", "va_lv_degree")
        for arr in (self.a_or, self.v_or, self.a_ex, self.v_ex, self.theta_or, self.theta_ex):
            arr[~np.isfinite(arr)] = 0.0
        self.v_or[~self.line_status] = 0.0
        self.v_ex[~self.line_status] = 0.0
        self.v_or[:] *= self.lines_or_pu_to_kv
        self.v_ex[:] *= self.lines_ex_pu_to_kv
        if self.n_storage > 0:
            sb = np.array([lab[int(b)] for b in net.storage.num("bus", 0)], dtype=np.int64)
            s_on = net.storage.flag("in_service")
            self.storage_p[:] = net.storage.num("p_mw", 0.0).astype(dt_float)
            self.storage_q[:] = net.storage.num("q_mvar", 0.0).astype(dt_float)
            sv = bus_vm[sb].astype(dt_float) * self.storage_pu_to_kv
            self.storage_theta[:] = sv          # sic: pPB:1635-1640 reads vm_pu for theta
            sv[~s_on] = 0.0
            self.storage_v[:] = sv
            dead = ~np.isfinite(self.storage_v)
            self.storage_p[dead] = 0.0
            self.storage_q[dead] = 0.0
            self.storage_v[dead] = 0.0
            for i in np.flatnonzero(dead):
                net.storage["in_service"][i] = False
        self.div_exception = None
        if is_dc:
            self.prod_q[:] = 0.0
            self.load_q[:] = 0.0
            self.storage_q[:] = 0.0
            self.q_or[:] = 0.0
            self.q_ex[:] = 0.0

    # pPB:1257-1287
    def _all_nan(self):
        for nm in ("p_or", "q_or", "v_or", "a_or", "p_ex", "q_ex", "v_ex", "a_ex", "prod_p", "prod_q", "prod_v",
                   "load_p", "load_q", "load_v", "storage_p", "storage_q", "storage_v", "theta_or", "theta_ex",
                   "load_theta", "gen_theta", "storage_theta"):
            getattr(self, nm)[:] = np.nan
        self._topo_vect = np.full_like(self._topo_vect, -1)
        self.line_status = np.zeros_like(self.line_status)

    # ------------------------------------------------------------------ getters (pPB:1439-1619)
    def get_line_status(self):
        return self.line_status

    def get_line_flow(self):
        return self.a_or

    def get_topo_vect(self):
        return self._topo_vect.copy()

    def generators_info(self):
        return self.prod_p.copy(), self.prod_q.copy(), self.prod_v.copy()

    def loads_info(self):
        return self.load_p.copy(), self.load_q.copy(), self.load_v.copy()

    def lines_or_info(self):
        return self.p_or.copy(), self.q_or.copy(), self.v_or.copy(), self.a_or.copy()

    def lines_ex_info(self):
        return self.p_ex.copy(), self.q_ex.copy(), self.v_ex.copy(), self.a_ex.copy()

    def storages_info(self):
        return self.storage_p.copy(), self.storage_q.copy(), self.storage_v.copy()

    def get_theta(self):
        return (1.0 * self.theta_or, 1.0 * self.theta_ex, 1.0 * self.load_theta, 1.0 * self.gen_theta,
                1.0 * self.storage_theta)

    def shunt_info(self):
        net = self._net
        lab = {int(l): i for i, l in enumerate(net.bus.index)}
        sb = np.array([lab[int(b)] for b in net.shunt.num("bus", 0)], dtype=np.int64)
        on = net.shunt.flag("in_service")
        p = net.res["shunt"]["p_mw"].astype(dt_float).copy()
        q = net.res["shunt"]["q_mvar"].astype(dt_float).copy()
        v = net.res["bus"]["vm_pu"][sb].astype(dt_float) * net.bus.num("vn_kv", 1.0)[sb].astype(dt_float)
        bus = self._g2l(net.shunt.num("bus", 0), on)
        v[~on] = 0.0
        return p, q, v, bus

    # ------------------------------------------------------------------ lifecycle
    def _disconnect_line(self, id_):   # pPB:1464-1475
        if id_ < self._n_true_line:
            self._net.line["in_service"][id_] = False
        else:
            self._net.trafo["in_service"][id_ - self._n_true_line] = False
        self._topo_vect[self.line_or_pos_topo_vect[id_]] = -1
        self._topo_vect[self.line_ex_pos_topo_vect[id_]] = -1
        self.line_status[id_] = False

    def reset(self, path=None, grid_filename=None) -> None:   # pPB:334-354
        self._net = self._net0.deepcopy()
        self._all_nan()
        self._refresh_status_topo()
        if self._net.res:
            self._fetch(is_dc=False)
        self.comp_time = 0.0

    def copy(self):   # pPB:1289-1409
        net, net0 = self._net, self._net0
        self._net = None
        self._net0 = None
        res = copy.deepcopy(self)
        self._net, self._net0 = net, net0
        res._net = net.deepcopy() if net is not None else None
        res._net0 = net0.deepcopy() if net0 is not None else None
        return res

    def close(self) -> None:
        self._net = None
        self._net0 = None

    def save_file(self, full_path) -> None:
        raise NotImplementedError("the oracle does not write grid files")
