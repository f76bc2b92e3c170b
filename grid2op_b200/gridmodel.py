"""Static grid description: pandapower-JSON tables -> what the device engine and the Backend need.

Follows the conventions of the reference backend so that element order, names, substation layout
and units are identical:

* element order / ``*_to_subid`` / ``*_to_sub_pos``     grid2op/Backend/pandaPowerBackend.py:670-769
* names                                                   pandaPowerBackend.py:481-553, 756-764
* reference-unit normalisation (ext_grid -> generator)    pandaPowerBackend.py:394-453
* thermal limits                                          pandaPowerBackend.py:806-813
* kV conversion vectors (float32)                         pandaPowerBackend.py:790-804

The electrical element model (per-unit pi equivalents of lines and two-winding transformers,
shunts) is the published pandapower model the reference delegates to (``pp.runpp`` at
pandaPowerBackend.py:1097); it is written here from scratch for the device data layout:
per line 4 complex admittances  yff, yft, ytf, ytt  so that  I_f = yff V_f + yft V_t.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .ppjson import PPNet, Table, read_pp_json

__all__ = ["GridModel"]


def _names(tbl: Table, prefix: str) -> np.ndarray:
    raw = tbl.strings("name")
    if len(tbl) and "name" in tbl and not tbl.has_nulls("name"):
        return np.array([str(v) for v in raw])
    bus = tbl.col("bus", 0).astype(np.int64)
    return np.array([f"{prefix}_{int(b)}_{i}" for i, b in enumerate(bus)])


class GridModel:
    """Immutable description shared by every copy of a backend (and by the batched driver)."""

    def __init__(self, path: str, n_busbar: int = 2):
        net = read_pp_json(path)
        self.path = path
        self.n_busbar = int(n_busbar)
        self.sn_mva = net.sn_mva
        self.f_hz = net.f_hz
        bus, line, trafo = net.table("bus"), net.table("line"), net.table("trafo")
        gen, eg = net.table("gen"), net.table("ext_grid")
        load, shunt, sto = net.table("load"), net.table("shunt"), net.table("storage")
        self.non_modeled = [k for k in ("trafo3w", "sgen", "switch", "motor", "asymmetric_load", "asymmetric_sgen",
                                        "impedance", "ward", "xward", "dcline", "measurement") if len(net.table(k))]
        self.n_sub = len(bus)
        labels = bus.index.astype(np.int64)
        if sorted(labels.tolist()) != list(range(self.n_sub)):
            raise ValueError("bus labels must be 0..n_sub-1 (same assumption as the reference, pPB:563)")
        vn_by_label = np.zeros(self.n_sub)
        vn_by_label[labels] = bus.col("vn_kv", 1.0)
        self.sub_vn = vn_by_label
        self.name_sub = np.array([f"sub_{int(i)}" for i in labels])
        self.bus_in_service0 = np.zeros(self.n_sub, dtype=bool)
        self.bus_in_service0[labels] = bus.col("in_service", 1.0) != 0
        self.bus_labels_file = labels.copy()                 # file row order (Backend.save_file)

        # ------------------------------------------------------------------ lines then trafos
        nl, nt = len(line), len(trafo)
        self.n_powerline, self.n_trafo = nl, nt
        self.n_line = nl + nt
        lf = line.col("from_bus", 0).astype(np.int64)
        lt = line.col("to_bus", 0).astype(np.int64)
        th = trafo.col("hv_bus", 0).astype(np.int64)
        tl = trafo.col("lv_bus", 0).astype(np.int64)
        self.line_or_sub = np.concatenate([lf, th]).astype(np.int32)
        self.line_ex_sub = np.concatenate([lt, tl]).astype(np.int32)
        self.line_in_service0 = np.concatenate([line.col("in_service", 1.0), trafo.col("in_service", 1.0)]) != 0
        if len(line) and "name" in line and not line.has_nulls("name"):
            names = [str(v) for v in line.strings("name")]
        else:
            names = [f"{int(a)}_{int(b)}_{i}" for i, (a, b) in enumerate(zip(lf, lt))]
        if nt and "name" in trafo and not trafo.has_nulls("name"):
            names += [str(v) for v in trafo.strings("name")]
        else:
            for i, (h, l) in enumerate(zip(th, tl)):
                a, b = sorted((str(int(h)), str(int(l))))      # string sort, like the reference
                names.append(f"{a}_{b}_{i + nl}")
        self.name_line = np.array(names)
        r, x, gch, bch, ratio, shift = self._branch_parameters(line, trafo, lf, th, tl)
        ys = 1.0 / (r + 1j * x)
        tap = ratio * np.exp(1j * np.deg2rad(shift))
        ytt = ys + 0.5 * (gch + 1j * bch)
        yff = ytt / (tap * np.conj(tap))
        yft = -ys / np.conj(tap)
        ytf = -ys / tap
        ly = np.zeros((self.n_line, 8))
        for k, y in enumerate((yff, yft, ytf, ytt)):
            ly[:, 2 * k] = y.real
            ly[:, 2 * k + 1] = y.imag
        self.line_y = ly
        self.line_bdc = 1.0 / x / ratio
        self.line_pshift = self.line_bdc * (-np.deg2rad(shift))
        self.line_or_vn = vn_by_label[self.line_or_sub].astype(np.float32)
        self.line_ex_vn = vn_by_label[self.line_ex_sub].astype(np.float32)
        self.thermal_limit_a = (1000.0 * np.concatenate([
            line.col("max_i_ka", 0.0),
            trafo.col("sn_mva", 0.0) / (np.sqrt(3.0) * trafo.col("vn_hv_kv", 1.0))])).astype(np.float32)

        # ------------------------------------------------------------------ generators / reference units
        g_bus = gen.col("bus", 0).astype(np.int64)
        g_p = gen.col("p_mw", 0.0) * gen.col("scaling", 1.0)
        g_vm = gen.col("vm_pu", 1.0)
        g_slack = gen.col("slack", 0.0) != 0
        g_qmin = gen.col("min_q_mvar", -1e9)
        g_qmax = gen.col("max_q_mvar", 1e9)
        g_on = gen.col("in_service", 1.0) != 0
        g_names: Optional[List[str]] = None
        if len(gen) and "name" in gen and not gen.has_nulls("name"):
            g_names = [str(v) for v in gen.strings("name")]
        e_bus = eg.col("bus", 0).astype(np.int64)
        self.ext_grid_bus_file = e_bus.copy()                # file row order (Backend.save_file)
        e_vm = eg.col("vm_pu", 1.0)
        e_qmin = eg.col("min_q_mvar", -1e9)
        e_qmax = eg.col("max_q_mvar", 1e9)
        e_on = eg.col("in_service", 1.0) != 0
        self.id_gen_added: Optional[int] = None     # index of the generator created on the ext_grid bus
        n_hidden = len(eg)
        if not g_slack.any():
            # pPB:394-451 - a generator is created where a reference unit sits on a bus without generator
            gen_buses = set(int(b) for b in g_bus)
            first = True
            for k in range(len(eg)):
                if int(e_bus[k]) in gen_buses:
                    continue
                g_bus = np.append(g_bus, e_bus[k]); g_p = np.append(g_p, 0.0); g_vm = np.append(g_vm, e_vm[k])
                g_slack = np.append(g_slack, first); g_qmin = np.append(g_qmin, e_qmin[k])
                g_qmax = np.append(g_qmax, e_qmax[k]); g_on = np.append(g_on, True)
                if g_names is not None:
                    g_names = None          # the created row has no name -> all names are regenerated
                if first:
                    self.id_gen_added = len(g_bus) - 1
                    first = False
                    n_hidden = min(n_hidden, 1)        # only the first ext_grid is kept (pPB:451)
        self.n_gen = len(g_bus)
        self.n_gen_file = len(gen)
        self.n_hidden = n_hidden
        self.gen_sub = g_bus.astype(np.int32)
        self.gen_p0, self.gen_vm0, self.gen_on0 = g_p.astype(np.float64), g_vm.astype(np.float64), g_on
        self.gen_slack = g_slack
        self.hidden_sub = e_bus[:n_hidden].astype(np.int32)
        self.hidden_vm0 = e_vm[:n_hidden].astype(np.float64)
        self.hidden_on0 = e_on[:n_hidden]
        self.name_gen = np.array(g_names) if g_names is not None else np.array(
            [f"gen_{int(b)}_{i}" for i, b in enumerate(g_bus)])
        self.unit_sub = np.concatenate([self.hidden_sub, self.gen_sub]).astype(np.int32)
        self.unit_is_ref = np.concatenate([np.ones(n_hidden, dtype=bool), g_slack]).astype(np.int32)
        self.unit_qmin = np.concatenate([e_qmin[:n_hidden], g_qmin]).astype(np.float64)
        self.unit_qmax = np.concatenate([e_qmax[:n_hidden], g_qmax]).astype(np.float64)
        self.unit_vn = vn_by_label[self.unit_sub].astype(np.float32)
        self.prod_pu_to_kv = vn_by_label[self.gen_sub].astype(np.float32)

        # ------------------------------------------------------------------ loads, storages, shunts
        self.n_load = len(load)
        self.load_sub = load.col("bus", 0).astype(np.int32)
        self.load_p0 = load.col("p_mw", 0.0) * load.col("scaling", 1.0)
        self.load_q0 = load.col("q_mvar", 0.0) * load.col("scaling", 1.0)
        self.load_on0 = load.col("in_service", 1.0) != 0
        self.name_load = _names(load, "load")
        self.load_vn = vn_by_label[self.load_sub].astype(np.float32)
        self.n_storage = len(sto)
        self.storage_sub = sto.col("bus", 0).astype(np.int32)
        self.storage_p0 = sto.col("p_mw", 0.0) * sto.col("scaling", 1.0)
        self.storage_q = sto.col("q_mvar", 0.0) * sto.col("scaling", 1.0)
        self.storage_on0 = sto.col("in_service", 1.0) != 0
        self.name_storage = _names(sto, "storage")
        self.storage_vn = vn_by_label[self.storage_sub].astype(np.float32)
        self.n_shunt = len(shunt)
        self.shunt_sub = shunt.col("bus", 0).astype(np.int32)
        self.shunt_p0 = shunt.col("p_mw", 0.0)
        self.shunt_q0 = shunt.col("q_mvar", 0.0)
        self.shunt_on0 = shunt.col("in_service", 1.0) != 0
        self.name_shunt = np.array([f"shunt_{int(b)}_{i}" for i, b in enumerate(self.shunt_sub)]).astype(str)
        self.shunt_vn = vn_by_label[self.shunt_sub].astype(np.float32)
        self.sh_vnkv = vn_by_label[self.shunt_sub].astype(np.float64)
        self.shunt_vratio = (vn_by_label[self.shunt_sub] / shunt.col("vn_kv", 1.0)) ** 2 * shunt.col("step", 1.0)

        # ------------------------------------------------------------------ substation layout (pPB:688-753)
        self.sub_info = np.zeros(self.n_sub, dtype=np.int32)
        used = np.zeros(self.n_sub, dtype=np.int32)

        def place(subs):
            pos = np.zeros(len(subs), dtype=np.int32)
            for i, s in enumerate(subs):
                pos[i] = used[s]
                used[s] += 1
                self.sub_info[s] += 1
            return pos

        self.line_or_to_sub_pos = np.zeros(self.n_line, dtype=np.int32)
        self.line_ex_to_sub_pos = np.zeros(self.n_line, dtype=np.int32)
        for i in range(self.n_line):           # origin then extremity, line by line
            so, se = int(self.line_or_sub[i]), int(self.line_ex_sub[i])
            self.line_or_to_sub_pos[i] = used[so]; used[so] += 1; self.sub_info[so] += 1
            self.line_ex_to_sub_pos[i] = used[se]; used[se] += 1; self.sub_info[se] += 1
        self.gen_to_sub_pos = place(self.gen_sub)
        self.load_to_sub_pos = place(self.load_sub)
        self.storage_to_sub_pos = place(self.storage_sub)
        self.dim_topo = int(self.sub_info.sum())
        off = np.concatenate([[0], np.cumsum(self.sub_info)[:-1]]).astype(np.int32)
        self.line_or_pos = (off[self.line_or_sub] + self.line_or_to_sub_pos).astype(np.int32)
        self.line_ex_pos = (off[self.line_ex_sub] + self.line_ex_to_sub_pos).astype(np.int32)
        self.gen_pos = (off[self.gen_sub] + self.gen_to_sub_pos).astype(np.int32)
        self.load_pos = (off[self.load_sub] + self.load_to_sub_pos).astype(np.int32)
        self.storage_pos = (off[self.storage_sub] + self.storage_to_sub_pos).astype(np.int32)
        self.unit_pos = np.concatenate([self.dim_topo + self.n_shunt + np.arange(n_hidden), self.gen_pos]).astype(np.int32)
        self.n_unit = self.n_hidden + self.n_gen
        self.sub_rank = self._bandwidth_order()
        self.n_slot = self.n_sub * self.n_busbar
        self.n_topo_in = self.dim_topo + self.n_shunt + self.n_hidden
        self.n_inj = self.n_gen + self.n_unit + 2 * self.n_load + self.n_storage + 2 * self.n_shunt
        self.n_out = 10 * self.n_line + 4 * self.n_unit + 2 * self.n_load + self.n_storage + 3 * self.n_shunt

    # ---------------------------------------------------------------------------------------------
    def _bandwidth_order(self) -> np.ndarray:
        """rank[sub] = position of the substation in a reverse Cuthill-McKee order of the substation graph:
        with buses numbered in that order the admittance / Jacobian matrices are banded (the
        CTA-per-instance kernel exploits it).  Deterministic; identity when scipy is unavailable."""
        n = self.n_sub
        try:
            from scipy.sparse import coo_matrix
            from scipy.sparse.csgraph import reverse_cuthill_mckee
        except Exception:  # pragma: no cover
            return np.arange(n, dtype=np.int32)
        if self.n_line == 0:
            return np.arange(n, dtype=np.int32)
        a, b = self.line_or_sub.astype(np.int64), self.line_ex_sub.astype(np.int64)
        m = coo_matrix((np.ones(2 * len(a)), (np.concatenate([a, b]), np.concatenate([b, a]))), shape=(n, n)).tocsr()
        order = np.asarray(reverse_cuthill_mckee(m, symmetric_mode=True), dtype=np.int64)
        rank = np.empty(n, dtype=np.int32)
        rank[order] = np.arange(n, dtype=np.int32)
        return rank

    # ---------------------------------------------------------------------------------------------
    def _branch_parameters(self, line: Table, trafo: Table, lf, th, tl):
        sn = self.sn_mva
        vn = self.sub_vn
        # lines: pi model on the origin-bus base
        length = line.col("length_km", 1.0)
        par = line.col("parallel", 1.0)
        zb = vn[lf] ** 2 / sn if len(lf) else np.zeros(0)
        r_l = line.col("r_ohm_per_km", 0.0) * length / par / zb
        x_l = line.col("x_ohm_per_km", 0.0) * length / par / zb
        b_l = 2.0 * np.pi * self.f_hz * line.col("c_nf_per_km", 0.0) * 1e-9 * length * par * zb
        g_l = line.col("g_us_per_km", 0.0) * 1e-6 * length * par * zb
        n_l = len(lf)
        # two-winding transformers ("t" model converted to pi, ratio tap changer on hv or lv side)
        n_t = len(th)
        r_t = np.zeros(n_t); x_t = np.zeros(n_t); g_t = np.zeros(n_t); b_t = np.zeros(n_t)
        ratio_t = np.ones(n_t)
        shift_t = trafo.col("shift_degree", 0.0)
        if n_t:
            vh = trafo.col("vn_hv_kv", 1.0).copy()
            vl = trafo.col("vn_lv_kv", 1.0).copy()
            vl_nom = trafo.col("vn_lv_kv", 1.0)
            dtap = trafo.col("tap_pos", np.nan) - trafo.col("tap_neutral", np.nan)
            step = trafo.col("tap_step_percent", np.nan)
            sides = trafo.strings("tap_side")
            for i in range(n_t):
                if np.isfinite(dtap[i]) and np.isfinite(step[i]):
                    k = 1.0 + dtap[i] * step[i] / 100.0
                    if sides[i] == "hv":
                        vh[i] *= k
                    elif sides[i] == "lv":
                        vl[i] *= k
            par_t = trafo.col("parallel", 1.0)
            sn_t = trafo.col("sn_mva", 1.0)
            vb_h, vb_l = vn[th], vn[tl]
            k_lv = (vl / vb_l) ** 2 * sn
            z = trafo.col("vk_percent", 0.0) / 100.0 / sn_t * k_lv
            rr = trafo.col("vkr_percent", 0.0) / 100.0 / sn_t * k_lv
            xx = np.sign(z) * np.sqrt(np.maximum(z * z - rr * rr, 0.0))
            r_t, x_t = rr / par_t, xx / par_t
            zb_l = vb_l ** 2 / sn
            pfe = trafo.col("pfe_kw", 0.0) * 1e-3
            i0 = trafo.col("i0_percent", 0.0)
            gm = pfe / vl_nom ** 2 * zb_l
            bm2 = np.maximum((i0 / 100.0 * sn_t) ** 2 - pfe ** 2, 0.0)
            bm = -np.sign(i0) * np.sqrt(bm2) * zb_l / vl_nom ** 2
            corr = (vl / vl_nom) ** 2
            gm, bm = gm / corr * par_t, bm / corr * par_t
            for i in range(n_t):
                if gm[i] == 0.0 and bm[i] == 0.0:
                    continue
                za = 0.5 * complex(r_t[i], x_t[i])
                zc = 1.0 / complex(gm[i], bm[i])
                zs = za * za + 2.0 * za * zc
                zab, zbc = zs / zc, zs / za
                r_t[i], x_t[i] = zab.real, zab.imag
                ysh = 2.0 / zbc
                g_t[i], b_t[i] = ysh.real, ysh.imag
            ratio_t = (vh / vl) / (vb_h / vb_l)
        r = np.concatenate([r_l, r_t]); x = np.concatenate([x_l, x_t])
        g = np.concatenate([g_l, g_t]); b = np.concatenate([b_l, b_t])
        ratio = np.concatenate([np.ones(n_l), ratio_t])
        shift = np.concatenate([np.zeros(n_l), shift_t])
        return r, x, g, b, ratio, shift

    # ---------------------------------------------------------------------------------------------
    def default_topo(self) -> np.ndarray:
        """int8 [n_topo_in]: everything that is in service in the file on busbar 1."""
        tv = np.full(self.n_topo_in, -1, dtype=np.int8)
        on = self.line_in_service0
        tv[self.line_or_pos[on]] = 1
        tv[self.line_ex_pos[on]] = 1
        tv[self.gen_pos[self.gen_on0]] = 1
        tv[self.load_pos[self.load_on0]] = 1
        if self.n_storage:
            tv[self.storage_pos[self.storage_on0]] = 1
        tv[self.dim_topo:self.dim_topo + self.n_shunt] = np.where(self.shunt_on0, 1, -1)
        tv[self.dim_topo + self.n_shunt:] = np.where(self.hidden_on0, 1, -1)
        return tv

    def default_inj(self) -> np.ndarray:
        """float64 [n_inj] with the set points stored in the grid file."""
        return np.concatenate([self.gen_p0, self.hidden_vm0, self.gen_vm0, self.load_p0, self.load_q0,
                               self.storage_p0, self.shunt_p0, self.shunt_q0]).astype(np.float64)

    def inj_slices(self):
        o = 0
        res = {}
        for name, n in (("gen_p", self.n_gen), ("hidden_vm", self.n_hidden), ("gen_vm", self.n_gen),
                        ("load_p", self.n_load), ("load_q", self.n_load), ("storage_p", self.n_storage),
                        ("shunt_p", self.n_shunt), ("shunt_q", self.n_shunt)):
            res[name] = slice(o, o + n)
            o += n
        assert o == self.n_inj
        return res

    # ---------------------------------------------------------------------------------------------
    # compact (de)serialisation: lets the batched driver / bench run without the grid2op data tree
    # ---------------------------------------------------------------------------------------------
    def to_npz(self, path: str) -> None:
        blob = {}
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                blob["a__" + k] = v
            elif isinstance(v, (int, float, str, bool)) or v is None:
                blob["s__" + k] = np.array([repr(v)])
            elif isinstance(v, list):
                blob["s__" + k] = np.array([repr(v)])
        np.savez_compressed(path, **blob)

    @classmethod
    def from_npz(cls, path: str) -> "GridModel":
        import ast

        self = object.__new__(cls)
        with np.load(path, allow_pickle=False) as z:
            for key in z.files:
                if key.startswith("a__"):
                    setattr(self, key[3:], z[key])
                else:
                    setattr(self, key[3:], ast.literal_eval(str(z[key][0])))
        return self
