"""``B200Backend`` — drop-in ``grid2op.Backend.Backend`` whose power flow runs on a B200.

Boundary mirrored (reference, abstract API): grid2op/Backend/backend.py:419-760
(``load_grid / apply_action / runpf / get_topo_vect / generators_info / loads_info /
lines_or_info / lines_ex_info`` + optional ``shunt_info / storages_info / get_theta / copy /
reset / _disconnect_line``).  Behavioural model: grid2op/Backend/pandaPowerBackend.py (cited per
method).  The host side keeps what PandaPowerBackend keeps in its pandas tables as flat numpy
arrays (set points in float64, per-element busbar + in-service flag); the solve is ONE call through
the C ABI (``include/b200pf.h``) into the CUDA engine.  There is no CPU solver in this package:
without ``libb200pf.so`` + a CUDA device ``load_grid`` raises.
"""
from __future__ import annotations

import copy
import os
import time
import warnings
from typing import Optional, Tuple, Union

import numpy as np

from ._bootstrap import ensure_grid2op, why_not

if not ensure_grid2op():  # pragma: no cover
    raise ImportError(f"grid2op (the host framework this backend plugs into) cannot be imported: {why_not()}")

from grid2op.Backend.backend import Backend  # noqa: E402
from grid2op.dtypes import dt_bool, dt_float, dt_int  # noqa: E402
from grid2op.Exceptions import BackendError  # noqa: E402

from .engine import ST_UNSUPPLIED, STATUS_TEXT, PowerFlowEngine  # noqa: E402
from .gridmodel import GridModel  # noqa: E402

__all__ = ["B200Backend"]

_K_LOAD, _K_GEN, _K_LOR, _K_LEX, _K_STO = 0, 1, 2, 3, 4


class B200Backend(Backend):
    shunts_data_available = True

    def __init__(self, detailed_infos_for_cascading_failures: bool = False, lightsim2grid: bool = False, dist_slack: bool = False,
                 max_iter: int = 10, can_be_copied: bool = True, with_numba: bool = False, *, tol_mva: float = 1e-8, device: int = 0):
        """Positional arguments as ``PandaPowerBackend.__init__`` (pPB:119-127) so that subclasses written for it keep working
        (reference tests/test_no_backend_copy.py:25-35).  ``lightsim2grid`` / ``with_numba`` select pandapower's inner solver
        there and mean nothing here; ``dist_slack`` (distributed slack) is not implemented."""
        if dist_slack:
            raise NotImplementedError("B200Backend: distributed slack (dist_slack=True) is not implemented")
        Backend.__init__(self, detailed_infos_for_cascading_failures=detailed_infos_for_cascading_failures,
                         can_be_copied=can_be_copied, lightsim2grid=lightsim2grid, dist_slack=dist_slack, max_iter=max_iter,
                         with_numba=with_numba, tol_mva=tol_mva, device=device)
        self._max_iter = int(max_iter)
        self._tol_mva = float(tol_mva)
        self._device = int(device)
        self.can_output_theta = True
        self._needs_active_bus = True                       # like pPB:144: apply_action reads the action's active_bus (bus.in_service of the reference)
        self._gm: Optional[GridModel] = None
        self._engine = None
        self._topo_vect: Optional[np.ndarray] = None
        self.div_exception = None
        self._last_status = 0
        self._last_iters = 0
        self._state0 = None

    # ------------------------------------------------------------------------------------------
    # engine factory (the only place that touches the device library)
    # ------------------------------------------------------------------------------------------
    def _make_engine(self, gm: GridModel):
        return PowerFlowEngine(gm, max_batch=1, device=self._device)

    # ------------------------------------------------------------------------------------------
    # load_grid                                                      reference pPB:356-617, 670-874
    # ------------------------------------------------------------------------------------------
    def load_grid(self, path: Union[os.PathLike, str], filename: Optional[Union[os.PathLike, str]] = None) -> None:
        self.can_handle_more_than_2_busbar()
        self.can_handle_detachment()
        full_path = self.make_complete_path(path, filename)
        gm = GridModel(full_path, n_busbar=self.n_busbar_per_sub)
        if gm.non_modeled:
            warnings.warn(f"There are {gm.non_modeled} in the grid file. These elements are not modeled "
                          "(neither by grid2op nor by this backend).")
        self._gm = gm
        self._engine = self._make_engine(gm)
        # --- grid description expected by grid2op (docs/grid2op_extend/createbackend.rst:322-347)
        self.n_line = gm.n_line
        self.n_gen = gm.n_gen
        self.n_load = gm.n_load
        self.n_sub = gm.n_sub
        self.name_line = gm.name_line.copy()
        self.name_gen = gm.name_gen.copy()
        self.name_load = gm.name_load.copy()
        self.name_sub = gm.name_sub.copy()
        self.n_storage = gm.n_storage
        if gm.n_storage == 0:
            self.set_no_storage()
        else:
            self.name_storage = gm.name_storage.copy()
            self.storage_to_subid = gm.storage_sub.astype(dt_int)
            self.storage_to_sub_pos = gm.storage_to_sub_pos.astype(dt_int)
        self.sub_info = gm.sub_info.astype(dt_int)
        self.dim_topo = int(gm.dim_topo)
        self.load_to_subid = gm.load_sub.astype(dt_int)
        self.gen_to_subid = gm.gen_sub.astype(dt_int)
        self.line_or_to_subid = gm.line_or_sub.astype(dt_int)
        self.line_ex_to_subid = gm.line_ex_sub.astype(dt_int)
        self.load_to_sub_pos = gm.load_to_sub_pos.astype(dt_int)
        self.gen_to_sub_pos = gm.gen_to_sub_pos.astype(dt_int)
        self.line_or_to_sub_pos = gm.line_or_to_sub_pos.astype(dt_int)
        self.line_ex_to_sub_pos = gm.line_ex_to_sub_pos.astype(dt_int)
        if type(self).shunts_data_available:                                    # pPB:556-559, 755-766
            self.n_shunt = gm.n_shunt
            self.shunt_to_subid = gm.shunt_sub.astype(dt_int)
            self.name_shunt = gm.name_shunt.copy()
            self._sh_vnkv = gm.sh_vnkv.copy()
        else:
            self.n_shunt = None         # (the shunts of the file still take part in the power flow, as set in the file)
        self._compute_pos_big_topo()
        self.thermal_limit_a = gm.thermal_limit_a.astype(dt_float)

        # the topology vector grid2op sees (normally the grid file's layout; without the storage units in a "compatibility"
        # environment, see storage_deact_for_backward_comaptibility)
        self._set_public_layout(gm.dim_topo, gm.line_or_pos, gm.line_ex_pos, gm.load_pos, gm.gen_pos,
                                gm.storage_pos if gm.n_storage else np.zeros(0, dtype=np.int64))

        # --- mutable state (what PandaPowerBackend keeps in its DataFrames)
        self._gen_p = gm.gen_p0.copy()
        self._gen_vm = gm.gen_vm0.copy()
        self._hid_vm = gm.hidden_vm0.copy()
        self._load_p = gm.load_p0.copy()
        self._load_q = gm.load_q0.copy()
        self._sto_p = gm.storage_p0.copy()
        self._sh_p = gm.shunt_p0.copy()
        self._sh_q = gm.shunt_q0.copy()
        # bus.in_service of the reference's net (pPB:920-922 rewrites it from the action's active_bus on every apply_action; the file
        # state — and so the state after reset() — has busbar 1 of every substation in service, the duplicated busbars out, pPB:548-566)
        self._active_bus = np.zeros((gm.n_sub, gm.n_busbar), dtype=bool)
        self._active_bus[:, 0] = True
        self._line_on = gm.line_in_service0.copy()
        self._lor_bus = np.ones(gm.n_line, dtype=np.int8)
        self._lex_bus = np.ones(gm.n_line, dtype=np.int8)
        self._gen_on = gm.gen_on0.copy()
        self._gen_bus = np.ones(gm.n_gen, dtype=np.int8)
        self._load_on = gm.load_on0.copy()
        self._load_bus = np.ones(gm.n_load, dtype=np.int8)
        self._sto_on = gm.storage_on0.copy()
        self._sto_bus = np.ones(gm.n_storage, dtype=np.int8)
        self._sh_on = gm.shunt_on0.copy()
        self._sh_bus = np.ones(gm.n_shunt, dtype=np.int8)
        self._hid_on = gm.hidden_on0.copy()
        self._hid_bus = np.ones(gm.n_hidden, dtype=np.int8)

        # --- result buffers                                                   pPB:815-860
        def buf(n):
            return np.full(n, np.nan, dtype=dt_float)

        self.p_or, self.q_or, self.v_or, self.a_or = buf(gm.n_line), buf(gm.n_line), buf(gm.n_line), buf(gm.n_line)
        self.p_ex, self.q_ex, self.v_ex, self.a_ex = buf(gm.n_line), buf(gm.n_line), buf(gm.n_line), buf(gm.n_line)
        self.theta_or, self.theta_ex = buf(gm.n_line), buf(gm.n_line)
        self.prod_p, self.prod_q, self.prod_v, self.gen_theta = buf(gm.n_gen), buf(gm.n_gen), buf(gm.n_gen), buf(gm.n_gen)
        self.load_p, self.load_q, self.load_v, self.load_theta = buf(gm.n_load), buf(gm.n_load), buf(gm.n_load), buf(gm.n_load)
        self.storage_p, self.storage_q, self.storage_v, self.storage_theta = (buf(gm.n_storage), buf(gm.n_storage),
                                                                              buf(gm.n_storage), buf(gm.n_storage))
        self._shunt_p, self._shunt_q, self._shunt_v = buf(gm.n_shunt), buf(gm.n_shunt), buf(gm.n_shunt)
        self.line_status = np.zeros(gm.n_line, dtype=dt_bool)
        self._compat_no_storage = False
        self._refresh_status_topo()
        self.comp_time = 0.0

        # initial power flow (the reference runs it when loading, pPB:386/455/577, AC then DC on failure)
        ok, _ = self.runpf(is_dc=False)
        if not ok:
            ok, _ = self.runpf(is_dc=True)
        if ok and gm.id_gen_added is not None:
            self._gen_p[gm.id_gen_added] = float(self.prod_p[gm.id_gen_added])   # pPB:422
        self._refresh_status_topo()
        self.comp_time = 0.0
        self._state0 = self._snapshot()

    # ------------------------------------------------------------------------------------------
    def _set_public_layout(self, dim_topo, p_lor, p_lex, p_load, p_gen, p_sto):
        """positions of the elements in the topology vector exchanged with grid2op + the position -> (kind, element)
        dispatch of apply_action (pPB:630-668)"""
        as_i = lambda a: np.asarray(a, dtype=np.int64)          # noqa: E731
        self._pub_dim = int(dim_topo)
        self._pub_lor, self._pub_lex, self._pub_load, self._pub_gen, self._pub_sto = as_i(p_lor), as_i(p_lex), as_i(p_load), as_i(p_gen), as_i(p_sto)
        kind = np.full(self._pub_dim, -1, dtype=np.int8)
        idx = np.zeros(self._pub_dim, dtype=np.int32)
        kind[self._pub_load] = _K_LOAD; idx[self._pub_load] = np.arange(len(self._pub_load))
        kind[self._pub_gen] = _K_GEN; idx[self._pub_gen] = np.arange(len(self._pub_gen))
        kind[self._pub_lor] = _K_LOR; idx[self._pub_lor] = np.arange(len(self._pub_lor))
        kind[self._pub_lex] = _K_LEX; idx[self._pub_lex] = np.arange(len(self._pub_lex))
        if len(self._pub_sto):
            kind[self._pub_sto] = _K_STO; idx[self._pub_sto] = np.arange(len(self._pub_sto))
        self._pos_kind, self._pos_idx = kind, idx

    def storage_deact_for_backward_comaptibility(self) -> None:                 # pPB:876-886 (called by Environment.py:690)
        """An environment in "compatibility" mode (``_compat_glop_version`` of a grid2op without storage units) has stripped
        the storage units from the class: results and topology vector take the class's sizes / positions.  The units stay in
        the power flow with the set points of the grid file, like in the reference's pandapower grid."""
        cls = type(self)
        self._compat_no_storage = cls.n_storage == 0 and self._gm.n_storage > 0
        if not self._compat_no_storage:
            return
        self._set_public_layout(cls.dim_topo, cls.line_or_pos_topo_vect, cls.line_ex_pos_topo_vect, cls.load_pos_topo_vect,
                                cls.gen_pos_topo_vect, np.zeros(0, dtype=np.int64))
        self.n_storage = 0
        self.dim_topo = int(cls.dim_topo)
        self._refresh_status_topo()

    def _snapshot(self):
        names = ("_active_bus", "_gen_p", "_gen_vm", "_hid_vm", "_load_p", "_load_q", "_sto_p", "_sh_p", "_sh_q", "_line_on",
                 "_lor_bus", "_lex_bus", "_gen_on", "_gen_bus", "_load_on", "_load_bus", "_sto_on", "_sto_bus",
                 "_sh_on", "_sh_bus", "_hid_on", "_hid_bus",
                 "p_or", "q_or", "v_or", "a_or", "p_ex", "q_ex", "v_ex", "a_ex", "theta_or", "theta_ex",
                 "prod_p", "prod_q", "prod_v", "gen_theta", "load_p", "load_q", "load_v", "load_theta",
                 "storage_p", "storage_q", "storage_v", "storage_theta", "_shunt_p", "_shunt_q", "_shunt_v")
        return {k: getattr(self, k).copy() for k in names}

    def _restore(self, snap):
        for k, v in snap.items():
            setattr(self, k, v.copy())

    # pPB:1450-1459 and 1489-1524
    def _refresh_status_topo(self):
        gm = self._gm
        self.line_status = self._line_on.astype(dt_bool)
        tv = np.full(self._pub_dim, -1, dtype=dt_int)
        tv[self._pub_lor] = np.where(self._line_on, self._lor_bus, -1)
        tv[self._pub_lex] = np.where(self._line_on, self._lex_bus, -1)
        tv[self._pub_load] = np.where(self._load_on, self._load_bus, -1)
        tv[self._pub_gen] = np.where(self._gen_on, self._gen_bus, -1)
        if len(self._pub_sto):
            tv[self._pub_sto] = np.where(self._sto_on, self._sto_bus, -1)
        self._topo_vect = tv

    # ------------------------------------------------------------------------------------------
    # apply_action                                                           reference pPB:902-1067
    # ------------------------------------------------------------------------------------------
    def apply_action(self, backend_action) -> None:
        if backend_action is None:
            return
        gm = self._gm
        active_bus, (prod_p, prod_v, load_p, load_q, storage), topo__, shunts__ = backend_action()
        ab = np.asarray(active_bus, dtype=bool)                                 # pPB:920-922
        if ab.shape == self._active_bus.shape:
            self._active_bus[:, :] = ab
        else:                                   # (an action class of another grid shape: no information -> the rule of runpf is off)
            self._active_bus[:, :] = False
        ch = prod_p.changed
        self._gen_p[ch] = prod_p.values[ch]
        ch = prod_v.changed
        # float32 / float32 like the reference (pPB:927); stored in float64
        self._gen_vm[ch] = prod_v.values[ch] / gm.prod_pu_to_kv[ch]
        if gm.id_gen_added is not None and gm.n_hidden and prod_v.changed[gm.id_gen_added]:
            self._hid_vm[:] = self._gen_vm[gm.id_gen_added]                    # pPB:929-931
        ch = load_p.changed
        self._load_p[ch] = load_p.values[ch]
        ch = load_q.changed
        self._load_q[ch] = load_q.values[ch]
        if gm.n_storage > 0 and not self._compat_no_storage:
            ch = storage.changed
            self._sto_p[ch] = storage.values[ch]
            stor_bus = backend_action.get_storages_bus()                        # pPB:941-951
            for i in np.flatnonzero(stor_bus.changed):
                b = int(stor_bus.values[i])
                if b <= -1:
                    self._sto_on[i] = False
                    self._sto_bus[i] = 1
                else:
                    self._sto_on[i] = True
                    self._sto_bus[i] = b
        if type(self).shunts_data_available and shunts__ is not None:
            shunt_p, shunt_q, shunt_bus = shunts__
            ch = shunt_p.changed
            self._sh_p[ch] = shunt_p.values[ch]
            ch = shunt_q.changed
            self._sh_q[ch] = shunt_q.values[ch]
            for i in np.flatnonzero(shunt_bus.changed):
                b = int(shunt_bus.values[i])
                if b == -1:
                    self._sh_on[i] = False
                else:
                    self._sh_on[i] = True
                    self._sh_bus[i] = b
        kind, idx = self._pos_kind, self._pos_idx
        for pos, new_bus in topo__:                                             # pPB:971-975
            k = kind[pos]
            i = idx[pos]
            new_bus = int(new_bus)
            if k == _K_LOAD:
                if new_bus >= 1:
                    self._load_on[i] = True; self._load_bus[i] = new_bus
                else:
                    self._load_on[i] = False
            elif k == _K_GEN:
                if new_bus >= 1:
                    self._gen_on[i] = True; self._gen_bus[i] = new_bus
                    if gm.id_gen_added is not None and gm.n_hidden and i == gm.n_gen - 1:
                        self._hid_bus[0] = new_bus                              # pPB:997-1003
                else:
                    self._gen_on[i] = False
            elif k == _K_LOR:
                if new_bus >= 1:
                    self._line_on[i] = True; self._lor_bus[i] = new_bus
                else:
                    self._line_on[i] = False
            elif k == _K_LEX:
                if new_bus >= 1:
                    self._line_on[i] = True; self._lex_bus[i] = new_bus
                else:
                    self._line_on[i] = False

    # ------------------------------------------------------------------------------------------
    # runpf                                                        reference pPB:1078-1120, 1220-1255
    # ------------------------------------------------------------------------------------------
    def _device_records(self):
        gm = self._gm
        topo = np.empty(gm.n_topo_in, dtype=np.int8)
        topo[gm.line_or_pos] = np.where(self._line_on, self._lor_bus, -1)
        topo[gm.line_ex_pos] = np.where(self._line_on, self._lex_bus, -1)
        topo[gm.load_pos] = np.where(self._load_on, self._load_bus, -1)
        topo[gm.gen_pos] = np.where(self._gen_on, self._gen_bus, -1)
        if gm.n_storage:
            topo[gm.storage_pos] = np.where(self._sto_on, self._sto_bus, -1)
        topo[gm.dim_topo:gm.dim_topo + gm.n_shunt] = np.where(self._sh_on, self._sh_bus, -1)
        topo[gm.dim_topo + gm.n_shunt:] = np.where(self._hid_on, self._hid_bus, -1)
        inj = np.concatenate([self._gen_p, self._hid_vm, self._gen_vm, self._load_p, self._load_q,
                              self._sto_p, self._sh_p, self._sh_q])
        return topo, inj

    def runpf(self, is_dc: bool = False) -> Tuple[bool, Union[Exception, None]]:
        gm = self._gm
        t0 = time.perf_counter()
        self._refresh_status_topo()                                             # pPB:1236-1237
        topo, inj = self._device_records()
        slots = np.unique(self._active_slots(topo))
        nb_cap = int(min(gm.n_slot, slots.size))                                   # exact number of active buses
        # A busbar the ENVIRONMENT holds for active (``active_bus``: it carries an element of the environment's topology vector or a
        # shunt the environment believes connected, _backendAction.py:1519-1531) is an in-service bus of the reference's net.  When no
        # element of the backend actually sits on it — a shunt that is out of service in the grid file while the environment counts it
        # in (l2rpn_neurips_2020_track1), once every other element of its substation has moved to the other busbar — pandapower leaves
        # that bus without a voltage and the reference reports a divergence (pPB:1241-1244 "Isolated bus").  Same here, before any
        # launch.
        has_el = np.zeros(gm.n_slot, dtype=bool)
        has_el[slots] = True
        if (self._active_bus.T.reshape(-1) & ~has_el).any():
            self.comp_time += time.perf_counter() - t0
            msg = STATUS_TEXT[ST_UNSUPPLIED]
            self._last_status, self._last_iters = ST_UNSUPPLIED, 0
            self.div_exception = BackendError(msg)
            self._reset_all_nan()
            return False, BackendError(f'powerflow diverged with error :"{msg}"')
        out, status, iters, _ = self._engine.run(topo[None, :], inj[None, :], is_dc=is_dc, max_iter=self._max_iter,
                                                 tol_mva=self._tol_mva, nb_cap=nb_cap)
        self.comp_time += time.perf_counter() - t0
        st = int(status[0])
        self._last_status, self._last_iters = st, int(iters[0])
        if st != 0:
            msg = STATUS_TEXT.get(st, f"status {st}")
            self.div_exception = BackendError(msg)
            self._reset_all_nan()
            return False, BackendError(f'powerflow diverged with error :"{msg}"')
        self._fetch(out, is_dc)
        self.div_exception = None
        return True, None

    def _active_slots(self, topo):
        gm = self._gm
        cache = getattr(gm, "_slot_index_cache", None)           # (constant per grid: substation and record position of every element)
        if cache is None:
            subs = np.concatenate([gm.line_or_sub, gm.line_ex_sub, gm.gen_sub, gm.load_sub, gm.storage_sub,
                                   gm.shunt_sub, gm.hidden_sub]).astype(np.int64)
            pos = np.concatenate([gm.line_or_pos, gm.line_ex_pos, gm.gen_pos, gm.load_pos, gm.storage_pos,
                                  gm.dim_topo + np.arange(gm.n_shunt), gm.dim_topo + gm.n_shunt + np.arange(gm.n_hidden)]).astype(np.int64)
            cache = gm._slot_index_cache = (subs, pos)
        subs, pos = cache
        b = topo[pos].astype(np.int64)
        ok = b > 0
        return subs[ok] + (b[ok] - 1) * gm.n_sub

    # pPB:1122-1218 (+ 1526-1564, 1596-1612, 1621-1647)
    def _fetch(self, out, is_dc):
        gm = self._gm
        v = self._engine.view(out)
        nh = gm.n_hidden
        self.p_or[:] = v.p_or[0]; self.q_or[:] = v.q_or[0]; self.v_or[:] = v.v_or[0]; self.a_or[:] = v.a_or[0]
        self.p_ex[:] = v.p_ex[0]; self.q_ex[:] = v.q_ex[0]; self.v_ex[:] = v.v_ex[0]; self.a_ex[:] = v.a_ex[0]
        self.theta_or[:] = v.theta_or[0]; self.theta_ex[:] = v.theta_ex[0]
        if not self._line_on.all():
            self._theta_of_open_lines(v)
        self.prod_p[:] = v.unit_p[0, nh:]; self.prod_q[:] = v.unit_q[0, nh:]
        self.prod_v[:] = v.unit_v[0, nh:]; self.gen_theta[:] = v.unit_theta[0, nh:]
        if gm.id_gen_added is not None and nh:
            # the created generator carries the power of the ext_grid it stands for     pPB:1536-1546
            self.prod_p[gm.id_gen_added] += v.unit_p[0, 0]
            self.prod_q[gm.id_gen_added] += v.unit_q[0, 0]
        self.load_p[:] = np.where(self._load_on, self._load_p, 0.0)
        self.load_q[:] = np.where(self._load_on, self._load_q, 0.0)
        self.load_v[:] = v.load_v[0]; self.load_theta[:] = v.load_theta[0]
        if gm.n_storage:
            self.storage_p[:] = self._sto_p                                    # set point table, pPB:1627
            self.storage_q[:] = gm.storage_q
            self.storage_v[:] = v.storage_v[0]
            self.storage_theta[:] = v.storage_v[0]                              # sic, pPB:1635-1640
            if not self._sto_on.all() and not is_dc:
                self._theta_of_detached_storages(v)
            if is_dc:
                # res_bus.vm_pu is NaN in DC -> the reference zeroes p, q, v        pPB:1203-1206
                # (only for connected units: a disconnected one already has v = 0, keeps its set point)
                on = self._sto_on.copy()
                self.storage_p[on] = 0.0; self.storage_q[on] = 0.0; self.storage_v[:] = 0.0
                self.storage_theta[:] = np.nan
                # ... and switches those units OFF in its grid (pPB:1207 `storage["in_service"].values[deact_storage] = False`):
                # this step's topology vector was taken before the solve (pPB:1236-1237) and still shows them connected, the next
                # one shows -1 and the units stay out until an action re-attaches them.  Reproduced as is.
                self._sto_on[on] = False
        self._shunt_p[:] = v.shunt_p[0]; self._shunt_q[:] = v.shunt_q[0]; self._shunt_v[:] = v.shunt_v[0]
        if is_dc:                                                               # pPB:1212-1218
            self.prod_q[:] = 0.0; self.load_q[:] = 0.0; self.storage_q[:] = 0.0
            self.q_or[:] = 0.0; self.q_ex[:] = 0.0
            # shunt_info reads res_bus.vm_pu (pPB:1599-1608), which pandapower leaves NaN after a DC solve: connected shunts report
            # v = NaN (disconnected ones 0, pPB:1610)
            self._shunt_v[self._sh_on] = np.nan

    def _theta_of_open_lines(self, v):
        """Reference quirk (pPB:1163-1187): the voltage of an out-of-service line / transformer is forced to 0 (pPB:1180-1181:
        "pandapower does not take into account disconnected powerline for their voltage") but its ANGLE is not — ``theta_or`` /
        ``theta_ex`` keep what pandapower's result table holds, the angle of the bus the end was last attached to
        (``line.from_bus`` is not touched by a disconnection, pPB:1040-1067), or 0 when that bus has no element left (NaN -> 0,
        pPB:1186-1187).  The kernels' result record carries 0 for open lines; the angle of the (substation, busbar) pair is the
        one any connected line end / generator / load on it reports (the same float32 value)."""
        gm = self._gm
        ns, nh = gm.n_sub, gm.n_hidden
        slot_theta = np.full(gm.n_slot, np.nan, dtype=np.float64)
        on = self._line_on
        slot_theta[gm.line_or_sub[on] + (self._lor_bus[on].astype(np.int64) - 1) * ns] = v.theta_or[0][on]
        slot_theta[gm.line_ex_sub[on] + (self._lex_bus[on].astype(np.int64) - 1) * ns] = v.theta_ex[0][on]
        g = self._gen_on
        slot_theta[gm.gen_sub[g] + (self._gen_bus[g].astype(np.int64) - 1) * ns] = v.unit_theta[0, nh:][g]
        ld = self._load_on
        slot_theta[gm.load_sub[ld] + (self._load_bus[ld].astype(np.int64) - 1) * ns] = v.load_theta[0][ld]
        off = ~on
        th_or = slot_theta[gm.line_or_sub[off] + (self._lor_bus[off].astype(np.int64) - 1) * ns]
        th_ex = slot_theta[gm.line_ex_sub[off] + (self._lex_bus[off].astype(np.int64) - 1) * ns]
        self.theta_or[off] = np.where(np.isfinite(th_or), th_or, 0.0)
        self.theta_ex[off] = np.where(np.isfinite(th_ex), th_ex, 0.0)

    def _theta_of_detached_storages(self, v):
        """The (sic) of pPB:1635-1640 goes one step further: ``storage_theta`` is ``res_bus.vm_pu`` of the bus in the storage TABLE times
        the storage's nominal voltage, and unlike ``storage_v`` it is not zeroed for units that are out of service (pPB:1641) — a detached
        unit reports the voltage magnitude [kV] of the busbar it was last attached to (NaN when that busbar has no element left)."""
        gm = self._gm
        ns, nh = gm.n_sub, gm.n_hidden
        slot_vm = np.full(gm.n_slot, np.nan, dtype=np.float64)
        on = self._line_on
        slot_vm[gm.line_or_sub[on] + (self._lor_bus[on].astype(np.int64) - 1) * ns] = v.v_or[0][on] / gm.line_or_vn[on]
        slot_vm[gm.line_ex_sub[on] + (self._lex_bus[on].astype(np.int64) - 1) * ns] = v.v_ex[0][on] / gm.line_ex_vn[on]
        g = self._gen_on
        slot_vm[gm.gen_sub[g] + (self._gen_bus[g].astype(np.int64) - 1) * ns] = v.unit_v[0, nh:][g] / gm.unit_vn[nh:][g]
        ld = self._load_on
        slot_vm[gm.load_sub[ld] + (self._load_bus[ld].astype(np.int64) - 1) * ns] = v.load_v[0][ld] / gm.load_vn[ld]
        st = self._sto_on
        slot_vm[gm.storage_sub[st] + (self._sto_bus[st].astype(np.int64) - 1) * ns] = v.storage_v[0][st] / gm.storage_vn[st]
        off = ~st
        self.storage_theta[off] = slot_vm[gm.storage_sub[off] + (self._sto_bus[off].astype(np.int64) - 1) * ns] * gm.storage_vn[off]

    # pPB:1257-1287
    def _reset_all_nan(self):
        for nm in ("p_or", "q_or", "v_or", "a_or", "p_ex", "q_ex", "v_ex", "a_ex", "prod_p", "prod_q", "prod_v",
                   "load_p", "load_q", "load_v", "storage_p", "storage_q", "storage_v", "theta_or", "theta_ex",
                   "load_theta", "gen_theta", "storage_theta", "_shunt_p", "_shunt_q", "_shunt_v"):
            getattr(self, nm)[:] = np.nan
        self._topo_vect = np.full(self._pub_dim, -1, dtype=dt_int)
        self.line_status = np.zeros(self._gm.n_line, dtype=dt_bool)

    # ------------------------------------------------------------------------------------------
    # read-back                                                              reference pPB:1439-1647
    # ------------------------------------------------------------------------------------------
    def get_line_status(self) -> np.ndarray:
        return self.line_status

    def get_line_flow(self) -> np.ndarray:
        return self.a_or

    def get_topo_vect(self) -> np.ndarray:
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
        if self._compat_no_storage:
            e = np.zeros(0, dtype=dt_float)
            return e, e.copy(), e.copy()
        return self.storage_p.copy(), self.storage_q.copy(), self.storage_v.copy()

    def shunt_info(self):
        bus = np.where(self._sh_on, self._sh_bus, -1).astype(dt_int)
        return self._shunt_p.copy(), self._shunt_q.copy(), self._shunt_v.copy(), bus

    def get_theta(self):
        return (1.0 * self.theta_or, 1.0 * self.theta_ex, 1.0 * self.load_theta, 1.0 * self.gen_theta,
                1.0 * self.storage_theta if not self._compat_no_storage else np.zeros(0, dtype=dt_float))

    def sub_from_bus_id(self, bus_id: int) -> int:
        n = self._gm.n_sub
        return bus_id - n * (bus_id // n)

    # ------------------------------------------------------------------------------------------
    # life cycle
    # ------------------------------------------------------------------------------------------
    def _disconnect_line(self, id_: int) -> None:                               # pPB:1464-1475
        self._line_on[id_] = False
        self._topo_vect[type(self).line_or_pos_topo_vect[id_]] = -1
        self._topo_vect[type(self).line_ex_pos_topo_vect[id_]] = -1
        self.line_status[id_] = False

    def reset(self, path=None, grid_filename=None) -> None:                     # pPB:334-354
        self._restore(self._state0)
        self._refresh_status_topo()
        self.comp_time = 0.0

    def copy(self) -> "B200Backend":                                            # pPB:1289-1409
        gm, eng, st0, kw = self._gm, self._engine, self._state0, self._my_kwargs
        self._gm = self._engine = self._state0 = None
        self._my_kwargs = None
        try:
            res = copy.deepcopy(self)
        finally:
            self._gm, self._engine, self._state0, self._my_kwargs = gm, eng, st0, kw
        res._gm, res._engine, res._state0 = gm, eng, st0      # immutable / shared device handle
        res._my_kwargs = kw         # the constructor arguments are handed on as they are, not copied (pPB:1296: type(self)(**self._my_kwargs))
        return res

    def close(self) -> None:
        self._engine = None
        self._gm = None

    def save_file(self, full_path) -> None:                                     # pPB:1425-1437 (pp.to_json)
        """Write the current grid state as a pandapower-JSON file: the source ``grid.json`` with the set points, element buses
        and in-service flags of this backend.  Like the reference's running net (``load_grid`` duplicates every bus once per extra
        busbar, pPB:548-566, and ``pp.to_json`` writes them), the bus table carries ``n_busbar_per_sub * n_sub`` rows: busbar k of
        substation s is bus ``s + (k-1) * n_sub``, in service when a connected element sits on it (pPB:920-922) — the file is a
        complete pandapower net of the CURRENT topology (tests/test_save_file.py solves it with the oracle's pandapower
        restatement and finds the backend's flows).  Loaded back as an environment grid it shows ``n_busbar_per_sub * n_sub``
        substations, exactly like a file saved by the reference."""
        from .ppjson import update_pp_json
        gm = self._gm
        nl, nf = gm.n_powerline, gm.n_gen_file

        def glob(sub, bar):
            return [int(s) + (int(b) - 1) * gm.n_sub for s, b in zip(sub, bar)]

        upd = {
            "line": {"in_service": list(self._line_on[:nl]), "from_bus": glob(gm.line_or_sub[:nl], self._lor_bus[:nl]),
                     "to_bus": glob(gm.line_ex_sub[:nl], self._lex_bus[:nl])},
            "load": {"p_mw": list(self._load_p), "q_mvar": list(self._load_q), "in_service": list(self._load_on),
                     "bus": glob(gm.load_sub, self._load_bus)},
            "gen": {"p_mw": list(self._gen_p[:nf]), "vm_pu": list(self._gen_vm[:nf]), "in_service": list(self._gen_on[:nf]),
                    "bus": glob(gm.gen_sub[:nf], self._gen_bus[:nf])},
        }
        if gm.n_trafo:
            upd["trafo"] = {"in_service": list(self._line_on[nl:]), "hv_bus": glob(gm.line_or_sub[nl:], self._lor_bus[nl:]),
                            "lv_bus": glob(gm.line_ex_sub[nl:], self._lex_bus[nl:])}
        if gm.n_shunt:
            upd["shunt"] = {"p_mw": list(self._sh_p), "q_mvar": list(self._sh_q), "in_service": list(self._sh_on),
                            "bus": glob(gm.shunt_sub, self._sh_bus)}
        if gm.n_storage:
            upd["storage"] = {"p_mw": list(self._sto_p), "in_service": list(self._sto_on), "bus": glob(gm.storage_sub, self._sto_bus)}
        if gm.n_hidden:
            upd["ext_grid"] = {"bus": glob(gm.hidden_sub, self._hid_bus) + [int(b) for b in gm.ext_grid_bus_file[gm.n_hidden:]]}
        # bus table: one row per (substation, busbar); active = a connected element sits on it
        active = np.zeros(gm.n_sub * gm.n_busbar, dtype=bool)
        for sub, bar, on in ((gm.line_or_sub, self._lor_bus, self._line_on), (gm.line_ex_sub, self._lex_bus, self._line_on),
                             (gm.load_sub, self._load_bus, self._load_on), (gm.gen_sub, self._gen_bus, self._gen_on),
                             (gm.storage_sub, self._sto_bus, self._sto_on), (gm.shunt_sub, self._sh_bus, self._sh_on),
                             (gm.hidden_sub, self._hid_bus, self._hid_on)):
            on = np.asarray(on, dtype=bool)
            if len(on):
                active[np.asarray(sub, dtype=np.int64)[on] + (np.asarray(bar, dtype=np.int64)[on] - 1) * gm.n_sub] = True
        row_of_label = {int(lab): r for r, lab in enumerate(gm.bus_labels_file)}
        dup = [(row_of_label[s], s + k * gm.n_sub) for k in range(1, gm.n_busbar) for s in sorted(row_of_label)]
        upd["bus"] = {"in_service": [bool(active[lab]) for lab in gm.bus_labels_file] + [bool(active[lab]) for _, lab in dup]}
        update_pp_json(gm.path, str(full_path), upd, duplicate_rows={"bus": dup})
