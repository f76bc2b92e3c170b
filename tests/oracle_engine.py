"""TEST INFRASTRUCTURE: an engine-shaped adapter around the fp64 oracle.

It has the call signature of ``grid2op_b200.engine.PowerFlowEngine.run`` (same packed records as
``include/b200pf.h``) but computes with ``oracle/pandapower_ref.py``.  Uses:
  * ``-m "not gpu"`` tests monkeypatch it into ``B200Backend._make_engine`` to exercise the HOST
    logic of the backend (apply_action, topology bookkeeping, read-back conventions) on the CPU;
  * gpu tests call it next to the CUDA engine on the same records and compare.
It is never importable from the product package.
"""
from __future__ import annotations

import numpy as np

from grid2op_b200.engine import OutputView
from grid2op_b200.gridmodel import GridModel
from oracle import pandapower_ref as ppr


class OracleEngine:
    def __init__(self, gm: GridModel, max_batch: int = 1, device: int = 0):
        self.gm = gm
        net = ppr.from_json(gm.path)
        # same reference-unit normalisation as the backend under test (pPB:394-453)
        if gm.id_gen_added is not None:
            n_file = len(net.gen)
            eg = net.ext_grid
            for k in range(gm.n_gen - n_file):
                net.gen.append_row((int(net.gen.index.max()) + 1) if len(net.gen) else 0,
                                   bus=int(gm.gen_sub[n_file + k]), p_mw=0.0, vm_pu=float(gm.gen_vm0[n_file + k]),
                                   min_q_mvar=float(gm.unit_qmin[gm.n_hidden + n_file + k]),
                                   max_q_mvar=float(gm.unit_qmax[gm.n_hidden + n_file + k]),
                                   scaling=1.0, slack=bool(gm.gen_slack[n_file + k]), in_service=True)
            net.tables["ext_grid"] = ppr.Frame(eg.index[:gm.n_hidden], {c: v[:gm.n_hidden] for c, v in eg.cols.items()})
        elif len(net.ext_grid) > gm.n_hidden:
            eg = net.ext_grid
            net.tables["ext_grid"] = ppr.Frame(eg.index[:gm.n_hidden], {c: v[:gm.n_hidden] for c, v in eg.cols.items()})
        if len(net.ext_grid):
            net.ext_grid["va_degree"] = np.array([0.0] * len(net.ext_grid), dtype=object)
        base = net.bus.copy()
        n_sub = gm.n_sub
        for k in range(1, gm.n_busbar):
            for i in range(n_sub):
                vals = {c: base.cols[c][i] for c in base.cols}
                vals["in_service"] = False
                net.bus.append_row(int(base.index[i]) + k * n_sub, **vals)
        self.net0 = net
        self.launch_count = 0

    def close(self):
        pass

    def view(self, out):
        return OutputView(self.gm, out)

    def last_launch_info(self):
        return {}

    def _one(self, tv, inj, is_dc, max_iter, tol_mva):
        gm = self.gm
        net = self.net0.deepcopy()
        sl = gm.inj_slices()
        n_sub = gm.n_sub

        def glob(sub, b):
            return int(sub) + (int(b) - 1) * n_sub

        active = np.zeros(gm.n_slot, dtype=bool)

        def setcol(tbl, col, i, val):
            tbl[col][i] = val

        nl = gm.n_powerline
        for l in range(gm.n_line):
            bo, be = int(tv[gm.line_or_pos[l]]), int(tv[gm.line_ex_pos[l]])
            on = bo > 0 and be > 0
            tbl, i = (net.line, l) if l < nl else (net.trafo, l - nl)
            cf, ct = ("from_bus", "to_bus") if l < nl else ("hv_bus", "lv_bus")
            setcol(tbl, "in_service", i, on)
            if on:
                setcol(tbl, cf, i, glob(gm.line_or_sub[l], bo)); setcol(tbl, ct, i, glob(gm.line_ex_sub[l], be))
            if bo > 0:
                active[glob(gm.line_or_sub[l], bo)] = True
            if be > 0:
                active[glob(gm.line_ex_sub[l], be)] = True
        for g in range(gm.n_gen):
            b = int(tv[gm.gen_pos[g]])
            setcol(net.gen, "in_service", g, b > 0)
            if b > 0:
                setcol(net.gen, "bus", g, glob(gm.gen_sub[g], b)); active[glob(gm.gen_sub[g], b)] = True
            setcol(net.gen, "p_mw", g, float(inj[sl["gen_p"]][g]))
            setcol(net.gen, "vm_pu", g, float(inj[sl["gen_vm"]][g]))
            if "scaling" in net.gen:
                setcol(net.gen, "scaling", g, 1.0)
        for h in range(gm.n_hidden):
            b = int(tv[gm.dim_topo + gm.n_shunt + h])
            setcol(net.ext_grid, "in_service", h, b > 0)
            if b > 0:
                setcol(net.ext_grid, "bus", h, glob(gm.hidden_sub[h], b)); active[glob(gm.hidden_sub[h], b)] = True
            setcol(net.ext_grid, "vm_pu", h, float(inj[sl["hidden_vm"]][h]))
        for k in range(gm.n_load):
            b = int(tv[gm.load_pos[k]])
            setcol(net.load, "in_service", k, b > 0)
            if b > 0:
                setcol(net.load, "bus", k, glob(gm.load_sub[k], b)); active[glob(gm.load_sub[k], b)] = True
            setcol(net.load, "p_mw", k, float(inj[sl["load_p"]][k])); setcol(net.load, "q_mvar", k, float(inj[sl["load_q"]][k]))
            if "scaling" in net.load:
                setcol(net.load, "scaling", k, 1.0)
        for k in range(gm.n_storage):
            b = int(tv[gm.storage_pos[k]])
            setcol(net.storage, "in_service", k, b > 0)
            if b > 0:
                setcol(net.storage, "bus", k, glob(gm.storage_sub[k], b)); active[glob(gm.storage_sub[k], b)] = True
            setcol(net.storage, "p_mw", k, float(inj[sl["storage_p"]][k]))
        for k in range(gm.n_shunt):
            b = int(tv[gm.dim_topo + k])
            setcol(net.shunt, "in_service", k, b > 0)
            if b > 0:
                setcol(net.shunt, "bus", k, glob(gm.shunt_sub[k], b)); active[glob(gm.shunt_sub[k], b)] = True
            setcol(net.shunt, "p_mw", k, float(inj[sl["shunt_p"]][k])); setcol(net.shunt, "q_mvar", k, float(inj[sl["shunt_q"]][k]))
        net.bus["in_service"] = np.array([bool(active[int(l)]) for l in net.bus.index], dtype=object)
        out = np.full(gm.n_out, np.nan, dtype=np.float32)
        busv = np.full(2 * gm.n_slot, np.nan)
        try:
            if is_dc:
                ppr.rundcpp(net, check_connectivity=True)
            else:
                ppr.runpp(net, max_iteration=max_iter, tolerance_mva=tol_mva, check_connectivity=False)
            bus_ok = net.bus.flag("in_service")
            if np.isnan(net.res["bus"]["va_degree"][bus_ok]).any():
                raise ppr.LoadflowNotConverged("unsupplied")
        except ppr.LoadflowNotConverged as exc:
            msg = str(exc)
            st = 2 if ("not connected" in msg or "unsupplied" in msg) else (3 if "no reference" in msg else 1)
            return out, st, 0, busv
        r = net.res
        v = OutputView(gm, out[None, :])
        lab = {int(l): i for i, l in enumerate(net.bus.index)}
        f32 = np.float32

        def cat(a, b):
            return np.concatenate([r["line"][a], r["trafo"][b]])

        on = np.array([tv[gm.line_or_pos[l]] > 0 and tv[gm.line_ex_pos[l]] > 0 for l in range(gm.n_line)])
        v.p_or[0] = cat("p_from_mw", "p_hv_mw"); v.q_or[0] = cat("q_from_mvar", "q_hv_mvar")
        v.p_ex[0] = cat("p_to_mw", "p_lv_mw"); v.q_ex[0] = cat("q_to_mvar", "q_lv_mvar")
        a1 = (cat("i_from_ka", "i_hv_ka") * 1000.0).astype(f32); a2 = (cat("i_to_ka", "i_lv_ka") * 1000.0).astype(f32)
        a1[~np.isfinite(a1)] = 0; a2[~np.isfinite(a2)] = 0
        v.a_or[0] = np.where(on, a1, 0); v.a_ex[0] = np.where(on, a2, 0)
        vm1 = cat("vm_from_pu", "vm_hv_pu").astype(f32); vm2 = cat("vm_to_pu", "vm_lv_pu").astype(f32)
        vm1[~np.isfinite(vm1)] = 0; vm2[~np.isfinite(vm2)] = 0
        v.v_or[0] = np.where(on, vm1 * gm.line_or_vn, 0); v.v_ex[0] = np.where(on, vm2 * gm.line_ex_vn, 0)
        t1 = cat("va_from_degree", "va_hv_degree").astype(f32); t2 = cat("va_to_degree", "va_lv_degree").astype(f32)
        t1[~np.isfinite(t1)] = 0; t2[~np.isfinite(t2)] = 0
        v.theta_or[0] = np.where(on, t1, 0); v.theta_ex[0] = np.where(on, t2, 0)
        nh = gm.n_hidden
        # ext_grid voltage = its bus voltage
        for h in range(nh):
            b = int(tv[gm.dim_topo + gm.n_shunt + h])
            if b > 0:
                i = lab[glob(gm.hidden_sub[h], b)]
                vmh = net.res["line"]["vm_from_pu"] if False else None
            v.unit_p[0, h] = r["ext_grid"]["p_mw"][h]; v.unit_q[0, h] = r["ext_grid"]["q_mvar"][h]
            if b > 0:
                pbus = ppr.pd2ppc(net)
                vm_b = (r["bus"]["vm_pu"][i] if not is_dc else pbus.vm0[i])
                v.unit_v[0, h] = f32(vm_b) * gm.unit_vn[h]; v.unit_theta[0, h] = r["bus"]["va_degree"][i]
            else:
                v.unit_v[0, h] = 0; v.unit_theta[0, h] = 0
        v.unit_p[0, nh:] = r["gen"]["p_mw"]; v.unit_q[0, nh:] = r["gen"]["q_mvar"]
        v.unit_v[0, nh:] = r["gen"]["vm_pu"].astype(f32) * gm.unit_vn[nh:]; v.unit_theta[0, nh:] = r["gen"]["va_degree"]
        vm_bus = r["bus"]["vm_pu"] if not is_dc else ppr.pd2ppc(net).vm0
        for k in range(gm.n_load):
            b = int(tv[gm.load_pos[k]])
            if b > 0:
                i = lab[glob(gm.load_sub[k], b)]
                v.load_v[0, k] = f32(vm_bus[i]) * gm.load_vn[k]; v.load_theta[0, k] = r["bus"]["va_degree"][i]
            else:
                v.load_v[0, k] = 0; v.load_theta[0, k] = 0
        for k in range(gm.n_storage):
            b = int(tv[gm.storage_pos[k]])
            v.storage_v[0, k] = f32(vm_bus[lab[glob(gm.storage_sub[k], b)]]) * gm.storage_vn[k] if b > 0 else 0
        for k in range(gm.n_shunt):
            b = int(tv[gm.dim_topo + k])
            v.shunt_p[0, k] = r["shunt"]["p_mw"][k]; v.shunt_q[0, k] = r["shunt"]["q_mvar"][k]
            v.shunt_v[0, k] = f32(vm_bus[lab[glob(gm.shunt_sub[k], b)]]) * gm.shunt_vn[k] if b > 0 else 0
        for s in range(gm.n_slot):
            i = lab[s]
            if active[s]:
                busv[s] = vm_bus[i]; busv[gm.n_slot + s] = np.deg2rad(r["bus"]["va_degree"][i])
        return out, 0, int(getattr(net, "iterations", 0)), busv

    def run(self, topo, inj, is_dc=False, max_iter=10, tol_mva=1e-8, nb_cap=0, want_busv=False):
        gm = self.gm
        topo = np.asarray(topo, dtype=np.int8).reshape(-1, gm.n_topo_in)
        inj = np.asarray(inj, dtype=np.float64).reshape(-1, gm.n_inj)
        B = topo.shape[0]
        out = np.empty((B, gm.n_out), dtype=np.float32)
        status = np.empty(B, dtype=np.int32)
        iters = np.empty(B, dtype=np.int32)
        busv = np.empty((B, 2 * gm.n_slot))
        for b in range(B):
            out[b], status[b], iters[b], busv[b] = self._one(topo[b], inj[b], is_dc, max_iter, tol_mva)
        self.launch_count += 1
        return out, status, iters, (busv if want_busv else None)


class COracleSeriesEngine:
    """TEST INFRASTRUCTURE: the series API of ``PowerFlowEngine`` (bind / set_topo / step / fetch) on top of the oracle's C
    restatement, WITHOUT protections — lets ``BatchedEnv`` / ``BatchedDoNothing`` host logic be checked on a machine without
    a GPU (``engine=`` injection).  Never importable from the product package."""

    def __init__(self, gm: GridModel):
        from oracle.c_oracle import COracle
        self.gm = gm
        self.orc = COracle(gm)
        self.prot = False
        self.launch_count = 0

    def series_bind(self, chron, scen, t0, static_inj, thermal_limit_a):
        self.chron = np.asarray(chron, dtype=np.float32)
        self.scen = np.asarray(scen, dtype=np.int64)
        self.t = np.asarray(t0, dtype=np.int64).copy()
        self.static_inj = np.asarray(static_inj, dtype=np.float64)
        self.th = np.asarray(thermal_limit_a, dtype=np.float32)
        self.B = len(self.scen)
        self.topo = np.tile(self.gm.default_topo(), (self.B, 1))

    def series_set_topo(self, topo):
        self.topo = np.ascontiguousarray(topo, dtype=np.int8).reshape(self.B, self.gm.n_topo_in).copy()

    def series_protections(self, enabled=True, *a, **k):
        if enabled:
            raise NotImplementedError("the CPU stand-in has no protections")

    def series_next_is_reset(self):
        pass

    def series_step(self, is_dc=False, max_iter=10, tol_mva=1e-8, nb_cap=0):
        gm = self.gm
        rows = self.chron[self.scen, self.t % self.chron.shape[1]]
        sl, nl, ng = gm.inj_slices(), gm.n_load, gm.n_gen
        inj = np.tile(self.static_inj, (self.B, 1))
        inj[:, sl["load_p"]] = rows[:, :nl]; inj[:, sl["load_q"]] = rows[:, nl:2 * nl]
        inj[:, sl["gen_p"]] = rows[:, 2 * nl:2 * nl + ng]
        inj[:, sl["gen_vm"]] = (rows[:, 2 * nl + ng:] / gm.prod_pu_to_kv[None, :]).astype(np.float32)
        self.out, self.status, self.iters, _ = self.orc.run(self.topo, inj, is_dc=is_dc, max_iter=max_iter, tol_mva=tol_mva)
        self.t = (self.t + 1) % self.chron.shape[1]
        self.launch_count += 1

    def series_fetch(self, want_out=True, want_rho=True):
        rho = OutputView(self.gm, self.out).a_or / self.th[None, :] if want_rho else None
        return (self.out if want_out else None), self.status, self.iters, rho

    def series_fetch_state(self):
        raise NotImplementedError

    def max_active_buses(self, topo, per_instance=False):
        """brute force (the product engine: b200pf_grid_max_active_buses)"""
        gm = self.gm
        subs = np.concatenate([gm.line_or_sub, gm.line_ex_sub, gm.gen_sub, gm.load_sub, gm.storage_sub, gm.shunt_sub, gm.hidden_sub])
        pos = np.concatenate([gm.line_or_pos, gm.line_ex_pos, gm.gen_pos, gm.load_pos, gm.storage_pos,
                              gm.dim_topo + np.arange(gm.n_shunt), gm.dim_topo + gm.n_shunt + np.arange(gm.n_hidden)])
        cnt = []
        for tv in np.asarray(topo).reshape(-1, gm.n_topo_in):
            b = tv[pos].astype(np.int64)
            cnt.append(len(set((subs[b > 0] * 8 + b[b > 0]).tolist())))
        best = max(cnt) if cnt else 0
        return (best, np.asarray(cnt, dtype=np.int32)) if per_instance else best

    def series_reset_instances(self, idx, t_new=None, topo_rows=None):
        idx = np.asarray(idx, dtype=np.int64)
        if t_new is not None:
            self.t[idx] = np.asarray(t_new, dtype=np.int64)
        if topo_rows is not None:
            self.topo[idx] = np.asarray(topo_rows, dtype=np.int8).reshape(len(idx), self.gm.n_topo_in)

    # rows entry point (what-if launches): pinned staging buffers are plain arrays here
    def staging(self):
        if not hasattr(self, "_stage"):
            gm = self.gm
            self._stage = dict(topo=np.zeros((self.B, gm.n_topo_in), dtype=np.int8), inj=np.zeros((self.B, gm.n_inj)),
                               out=np.zeros((self.B, gm.n_out), dtype=np.float32), status=np.zeros(self.B, dtype=np.int32),
                               iters=np.zeros(self.B, dtype=np.int32))
            self._rows = np.zeros((self.B, 2 * gm.n_load + 2 * gm.n_gen), dtype=np.float32)
        return self._stage

    def rows_staging(self):
        self.staging()
        return self._rows

    def set_static_inj(self, static_inj):
        self.static_inj = np.asarray(static_inj, dtype=np.float64)

    def run_rows_staged(self, batch, is_dc=False, max_iter=10, tol_mva=1e-8, nb_cap=0):
        gm = self.gm
        rows = self._rows[:batch]
        sl, nl, ng = gm.inj_slices(), gm.n_load, gm.n_gen
        inj = np.tile(self.static_inj, (batch, 1))
        inj[:, sl["load_p"]] = rows[:, :nl]; inj[:, sl["load_q"]] = rows[:, nl:2 * nl]
        inj[:, sl["gen_p"]] = rows[:, 2 * nl:2 * nl + ng]
        inj[:, sl["gen_vm"]] = (rows[:, 2 * nl + ng:] / gm.prod_pu_to_kv[None, :]).astype(np.float32)
        out, status, iters, _ = self.orc.run(self._stage["topo"][:batch], inj, is_dc=is_dc, max_iter=max_iter, tol_mva=tol_mva)
        self._stage["out"][:batch] = out; self._stage["status"][:batch] = status; self._stage["iters"][:batch] = iters

    def view(self, out):
        return OutputView(self.gm, out)

    def plan_stats(self):
        return {"last_kernel": "oracle"}

    def close(self):
        pass


class EmuProtSeriesEngine(COracleSeriesEngine):
    """TEST INFRASTRUCTURE: the series API WITH protections, on the host build of the planned kernel + its cascade loop
    (tests/emu/sparse_emu.cpp ``sparse_emu_series_prot_step``: the same kernel source and the same host loop as
    ``b200pf_series_step`` with protections on the 36 / 118-substation grids).  Lets ``BatchedEnv`` with protections be stepped next to
    unmodified environments on a machine without a GPU."""

    def __init__(self, gm: GridModel):
        super().__init__(gm)
        import ctypes as C
        from sparse_emu import SparseEmu
        self._C = C
        self.emu = SparseEmu(gm)
        self.emu.lib.sparse_emu_series_prot_step.restype = C.c_int
        self.next_reset = 0
        self.hard, self.soft, self.max_allowed = 2.0, 1.0, 2

    def series_bind(self, chron, scen, t0, static_inj, thermal_limit_a):
        super().series_bind(chron, scen, t0, static_inj, thermal_limit_a)
        gm, B, nl = self.gm, self.B, self.gm.n_line
        self.chron32 = np.ascontiguousarray(self.chron, dtype=np.float32)
        self.scen32 = np.ascontiguousarray(self.scen, dtype=np.int32)
        self.t32 = np.ascontiguousarray(self.t, dtype=np.int32).copy()
        self.topo = np.ascontiguousarray(self.topo, dtype=np.int8)
        self.out = np.zeros((B, gm.n_out), dtype=np.float32)
        self.status = np.zeros(B, dtype=np.int32); self.iters = np.zeros(B, dtype=np.int32)
        self.rho = np.zeros((B, nl), dtype=np.float32)
        self.pcount = np.zeros((B, nl), dtype=np.int32); self.ts_over = np.zeros((B, nl), dtype=np.int32)
        self.disc = np.full((B, nl), -1, dtype=np.int32); self.done = np.zeros(B, dtype=np.int32)
        self.rounds = np.zeros(1, dtype=np.int32)

    def series_set_topo(self, topo):
        self.topo = np.ascontiguousarray(topo, dtype=np.int8).reshape(self.B, self.gm.n_topo_in).copy()

    def series_protections(self, enabled=True, hard_overflow_threshold=2.0, soft_overflow_threshold=1.0, nb_timestep_overflow_allowed=2):
        self.prot = bool(enabled)
        self.hard, self.soft, self.max_allowed = float(hard_overflow_threshold), float(soft_overflow_threshold), int(nb_timestep_overflow_allowed)

    def series_next_is_reset(self):
        self.next_reset = 1

    def series_step(self, is_dc=False, max_iter=10, tol_mva=1e-8, nb_cap=0):
        if not self.prot or is_dc:
            self.t = self.t32.astype(np.int64)
            super().series_step(is_dc=is_dc, max_iter=max_iter, tol_mva=tol_mva, nb_cap=nb_cap)
            self.t32 = self.t.astype(np.int32)
            self.rho = (OutputView(self.gm, self.out).a_or / self.th[None, :]).astype(np.float32)
            return
        C = self._C
        p = lambda a: a.ctypes.data_as(C.c_void_p)     # noqa: E731
        static = np.ascontiguousarray(self.static_inj, dtype=np.float64)
        th = np.ascontiguousarray(self.th, dtype=np.float32)
        rc = self.emu.lib.sparse_emu_series_prot_step(
            C.byref(self.emu.desc), C.c_int(self.B), p(self.topo), p(self.chron32), C.c_int(self.chron32.shape[0]), C.c_int(self.chron32.shape[1]),
            p(self.scen32), p(self.t32), p(static), p(th), C.c_int(int(self.next_reset)), C.c_float(self.hard), C.c_float(self.soft),
            C.c_int(self.max_allowed), C.c_int(int(max_iter)), C.c_double(float(tol_mva)), p(self.out), p(self.status), p(self.iters), p(self.rho),
            p(self.pcount), p(self.ts_over), p(self.disc), p(self.done), p(self.rounds))
        assert rc == 0, rc
        self.next_reset = 0
        self.launch_count += 1

    def series_fetch(self, want_out=True, want_rho=True):
        return (self.out if want_out else None), self.status, self.iters, (self.rho if want_rho else None)

    def series_fetch_state(self):
        return dict(protection_counter=self.pcount.copy(), timestep_overflow=self.ts_over.copy(), disc_lines=self.disc.copy(), done=self.done.copy())

    def series_reset_instances(self, idx, t_new=None, topo_rows=None):
        idx = np.asarray(idx, dtype=np.int64)
        if t_new is not None:
            self.t32[idx] = np.asarray(t_new, dtype=np.int32)
        if topo_rows is not None:
            self.topo[idx] = np.asarray(topo_rows, dtype=np.int8).reshape(len(idx), self.gm.n_topo_in)
        self.pcount[idx] = 0; self.ts_over[idx] = 0; self.disc[idx] = -1; self.done[idx] = 0
