"""Batched ``env.step`` WITH actions: the driver next to :class:`grid2op_b200.rollout.BatchedDoNothing` for agents that
act on the topology (BASELINE.json configs[2]: "random topology actions"), with the environment-side bookkeeping a
topology agent depends on.

What one :meth:`BatchedEnv.step` restates for a whole batch of independent environments (reference, per env —
``BaseEnv.step`` grid2op/Environment/baseEnv.py:3778-3872 and what it calls), for the action subset
{do nothing, set the busbars of ONE substation, set the status of ONE line} (``MAX_SUB_CHANGED = MAX_LINE_STATUS_CHANGED
= 1``, grid2op/Parameters.py:279):

1. legality — ``DefaultRules``: an action that touches a line whose ``times_before_line_status_actionable`` is > 0 or a
   substation whose ``times_before_topology_actionable`` is > 0 is illegal and replaced by "do nothing"
   (grid2op/Rules/PreventReconnection.py:33-60, baseEnv.py:3798-3820);
2. ``_backend_action += action`` (baseEnv.py:3822-3846): busbars of the substation's elements, line disconnection
   (both ends -1) / reconnection (each end back on its last busbar, grid2op/Action/_backendAction.py ``last_topo_registered``);
   a busbar > 0 for the end of a DISCONNECTED line inside a substation action reconnects that line (the line, not the
   substation, is what such an entry impacts: grid2op/Action/baseAction.py:1836-1860);
3. the environment's own modifications of this step: next chronics row (gridStateFromFile.py:766-776), lines in
   ``maintenance`` / ``hazards`` forced out (gridStateFromFile.py:780-797, baseAction ``maintenance`` / ``hazards`` keys);
4. ``backend.next_grid_state`` — one batched power flow with the protections of backend.py:1433-1521 on the device;
5. the cooldown bookkeeping of baseEnv.py:3352-3393: line cooldowns decrease, lines tripped by the protections get
   ``NB_TIMESTEP_RECONNECTION``, maintenance / hazard durations are folded in (baseEnv.py:2566-2597), lines / substations
   the (legal) action named get ``NB_TIMESTEP_COOLDOWN_LINE`` / ``NB_TIMESTEP_COOLDOWN_SUB``;
6. game over when the power flow fails (baseEnv.py:3523-3524); a finished instance stays finished.

The topology lives on the HOST (int8 [batch, n_topo_in]); a step that changed it hands the rows to
``PowerFlowEngine.series_set_topo`` — cached plans are looked up by hash, new topologies are planned on the host threads
(see DESIGN.md "config 3").  Everything numeric stays on the device as in :class:`BatchedDoNothing`.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .gridmodel import GridModel
from .rollout import BatchedDoNothing

__all__ = ["BatchedEnv", "random_substation_actions"]


class BatchedEnv(BatchedDoNothing):
    def __init__(self, gm: GridModel, chron: np.ndarray, batch: int, *, maintenance: Optional[np.ndarray] = None,
                 hazards: Optional[np.ndarray] = None, nb_timestep_cooldown_line: int = 0, nb_timestep_cooldown_sub: int = 0,
                 nb_timestep_reconnection: int = 10, engine=None, tight_cap: bool = True, **kw):
        """``maintenance`` / ``hazards``: bool [n_scen, n_rows, n_line] (:func:`grid2op_b200.chronics.load_line_events` per
        scenario), row-aligned with ``chron``.  Cooldown parameters: grid2op.Parameters NB_TIMESTEP_COOLDOWN_LINE /
        NB_TIMESTEP_COOLDOWN_SUB / NB_TIMESTEP_RECONNECTION (Parameters.py:255-300)."""
        super().__init__(gm, chron, batch, engine=engine, **kw)
        B, nl = self.batch, gm.n_line
        self.topo = self.topo0.copy()                                      # int8 [B, n_topo_in], -1 = disconnected
        self.last_bus = np.where(self.topo[:, :gm.dim_topo] > 0, self.topo[:, :gm.dim_topo], 1).astype(np.int8)
        self.line_cooldown = np.zeros((B, nl), dtype=np.int32)             # times_before_line_status_actionable
        self.sub_cooldown = np.zeros((B, gm.n_sub), dtype=np.int32)        # times_before_topology_actionable
        self.done = np.zeros(B, dtype=bool)
        self.row = self.t0.astype(np.int64).copy()                         # chronics row the NEXT solve uses
        self.cd_line, self.cd_sub, self.cd_reco = int(nb_timestep_cooldown_line), int(nb_timestep_cooldown_sub), int(nb_timestep_reconnection)
        self._maint = self._maint_time = self._maint_dur = self._haz = self._haz_dur = None
        if maintenance is not None:
            from .chronics import maintenance_time_duration
            self._maint = np.asarray(maintenance, dtype=bool)
            td = [maintenance_time_duration(m) for m in self._maint]
            self._maint_time = np.stack([x[0] for x in td]); self._maint_dur = np.stack([x[1] for x in td])
        if hazards is not None:
            from .chronics import hazard_duration
            self._haz = np.asarray(hazards, dtype=bool)
            self._haz_dur = np.stack([hazard_duration(h) for h in self._haz])
        # element positions of every substation in the topology vector (grid2op order), padded
        sub_of = np.full(gm.dim_topo, -1, dtype=np.int64)
        sub_of[gm.line_or_pos] = gm.line_or_sub; sub_of[gm.line_ex_pos] = gm.line_ex_sub
        sub_of[gm.gen_pos] = gm.gen_sub; sub_of[gm.load_pos] = gm.load_sub
        if gm.n_storage:
            sub_of[gm.storage_pos] = gm.storage_sub
        self.sub_of_pos = sub_of
        self.sub_size = np.bincount(sub_of, minlength=gm.n_sub)
        self.max_sub_size = int(self.sub_size.max())
        self.sub_pos = np.full((gm.n_sub, self.max_sub_size), -1, dtype=np.int64)
        for s in range(gm.n_sub):
            p = np.flatnonzero(sub_of == s)
            self.sub_pos[s, :len(p)] = p
        self.line_or_pos, self.line_ex_pos = np.asarray(gm.line_or_pos, dtype=np.int64), np.asarray(gm.line_ex_pos, dtype=np.int64)
        # line whose end sits at a position of the topology vector (-1: not a line end) and the position of its other end
        self.line_of_pos = np.full(gm.dim_topo, -1, dtype=np.int64)
        self.other_end_pos = np.full(gm.dim_topo, -1, dtype=np.int64)
        self.line_of_pos[self.line_or_pos] = np.arange(nl); self.line_of_pos[self.line_ex_pos] = np.arange(nl)
        self.other_end_pos[self.line_or_pos] = self.line_ex_pos; self.other_end_pos[self.line_ex_pos] = self.line_or_pos
        self._topo_dirty = False
        # Shunts: this driver never acts on them, so the environment of the reference holds every shunt for connected to busbar 1
        # (_BackendAction.current_shunt_bus, grid2op/Action/_backendAction.py:513-514) and counts that busbar as active
        # (:1519-1531 -> bus.in_service, pandaPowerBackend.py:920-922) — also when the grid file has the shunt out of service (all six of
        # l2rpn_neurips_2020_track1).  Busbar 1 of such a substation must keep an element: :meth:`_isolated_shunt_busbar`.
        self._shunt_off = np.flatnonzero(self.topo0[0, gm.dim_topo:gm.dim_topo + gm.n_shunt] <= 0) if gm.n_shunt else np.zeros(0, dtype=np.int64)
        self._shunt_off_pos = [self.sub_pos[int(gm.shunt_sub[k]), :self.sub_size[int(gm.shunt_sub[k])]] for k in self._shunt_off]
        hid_sub = np.asarray(gm.hidden_sub, dtype=np.int64) if gm.n_hidden else np.zeros(0, dtype=np.int64)
        on_sub = np.asarray(gm.shunt_sub, dtype=np.int64)[np.setdiff1d(np.arange(gm.n_shunt), self._shunt_off)] if gm.n_shunt else hid_sub
        self._iso_cols = self._iso_seg = None
        self._shunt_off_other = [bool(np.isin(int(gm.shunt_sub[k]), hid_sub) or np.isin(int(gm.shunt_sub[k]), on_sub)) for k in self._shunt_off]
        # Bus splits add active buses: the kernels that discover the topology on the device size their workspace (and the
        # threads per instance) for ``nb_cap`` active buses.  The host knows every instance's topology, so the cap is the
        # bus count of the fullest instance (:meth:`_bus_cap`: host code of the library, recomputed when a topology changed) instead of every bus slot.
        self.tight_cap = bool(tight_cap)
        self._nb_inst = np.zeros(B, dtype=np.int32)                        # active buses per instance (kept incrementally)
        self._cap_dirty = np.ones(B, dtype=bool)                           # instances whose topology changed since it was counted
        self.nb_cap = self._bus_cap() if self.tight_cap else 0
        self.n_illegal = 0
        self.n_steps = 0

    # ------------------------------------------------------------------------------------------------------------
    def _bus_cap(self) -> int:
        """Active buses (bus slots with a connected element) of the fullest instance, rounded up to a multiple of 8 above 17
        (<= 17: exact — the warp-per-instance kernel takes systems up to 17 buses), at most every bus slot.  Per-instance counts are
        kept; only the instances whose topology changed since the last call are counted again (lines tripped by the protections
        only remove buses: the kept count stays an upper bound)."""
        rows = np.flatnonzero(self._cap_dirty)
        if len(rows) == self.batch:
            self._nb_inst[:] = self.engine.max_active_buses(self.topo, per_instance=True)[1]
        elif len(rows):                   # only the instances whose topology changed are counted again
            self._nb_inst[rows] = self.engine.max_active_buses(self.topo[rows], per_instance=True)[1]
        self._cap_dirty[:] = False
        n = int(self._nb_inst.max())
        if n > 17:
            n = (n + 7) // 8 * 8
        return min(n, self.gm.n_slot)

    def line_status(self) -> np.ndarray:
        return (self.topo[:, self.line_or_pos] > 0) & (self.topo[:, self.line_ex_pos] > 0)

    def _apply_agent(self, sub_id, sub_bus, line_id, line_status):
        """-> (aff_lines bool [B, n_line], aff_subs bool [B, n_sub]) of the LEGAL actions, applied to self.topo"""
        gm, B = self.gm, self.batch
        aff_l = np.zeros((B, gm.n_line), dtype=bool)
        aff_s = np.zeros((B, gm.n_sub), dtype=bool)
        idx = np.arange(B)
        has_sub = np.zeros(B, dtype=bool)
        has_line = np.zeros(B, dtype=bool)
        if sub_id is not None:
            sub_id = np.asarray(sub_id, dtype=np.int64)
            sub_bus = np.asarray(sub_bus, dtype=np.int8)
            has_sub = (sub_id >= 0) & (np.abs(sub_bus) > 0).any(axis=1)
        if line_id is not None:
            line_id = np.asarray(line_id, dtype=np.int64)
            line_status = np.asarray(line_status, dtype=np.int64)
            has_line = (line_id >= 0) & (line_status != 0)
        # A busbar > 0 for the end of a DISCONNECTED line is a reconnection of that line (reference: BaseAction.get_topological_impact,
        # grid2op/Action/baseAction.py:1836-1860 — the line counts as impacted, its two ends do not count for the substation; applied by
        # _BackendAction: the other end goes back to its last busbar)
        reco = None
        sub_impacted = has_sub.copy()
        n_reco = np.zeros(B, dtype=np.int64)
        reco_in_cd = np.zeros(B, dtype=bool)
        reco_is_line = np.zeros(B, dtype=bool)           # the explicitly named line is one of the reconnected ones
        if has_sub.any():
            who_s = np.flatnonzero(has_sub)
            pos = self.sub_pos[sub_id[who_s]]                                    # [n, max_sub_size]
            val = sub_bus[who_s][:, :self.max_sub_size]
            m = (pos >= 0) & (val > 0)
            psafe = np.where(pos >= 0, pos, 0)
            assert self.topo.flags.c_contiguous and self.last_bus.flags.c_contiguous
            tflat, lflat = self.topo.reshape(-1), self.last_bus.reshape(-1)       # (views: both arrays are C-contiguous)
            ft = who_s[:, None] * self.topo.shape[1] + psafe                      # flat indices of the named entries
            cur = tflat[ft]
            dead = m & (cur <= 0)
            if dead.any():                                                        # rare: some named element is disconnected
                lp = np.where(pos >= 0, self.line_of_pos[psafe], -1)
                reco = dead & (lp >= 0)
                if reco.any():
                    lsafe = np.where(lp >= 0, lp, 0)
                    n_reco[who_s] = reco.sum(axis=1)
                    reco_in_cd[who_s] = (reco & (self.line_cooldown[who_s[:, None], lsafe] > 0)).any(axis=1)
                    if line_id is not None:
                        reco_is_line[who_s] = (reco & (lp == np.where(has_line[who_s], line_id[who_s], -2)[:, None])).any(axis=1)
                    sub_impacted[who_s] = (m & ~reco).any(axis=1)
                else:
                    reco = None
        # legality (DefaultRules: LookParam MAX_SUB_CHANGED = MAX_LINE_STATUS_CHANGED = 1, grid2op/Rules/LookParam.py; PreventReconnection on
        # what the action impacts, grid2op/Rules/PreventReconnection.py:33-60); finished instances ignore their action
        illegal = np.zeros(B, dtype=bool)
        if has_sub.any():
            illegal |= sub_impacted & (self.sub_cooldown[idx, np.where(has_sub, sub_id, 0)] > 0)
            if reco is not None:
                illegal |= reco_in_cd
                illegal |= (n_reco + (has_line & ~reco_is_line)) > 1
        if has_line.any():
            illegal |= has_line & (self.line_cooldown[idx, np.where(has_line, line_id, 0)] > 0)
        self.n_illegal += int((illegal & ~self.done).sum())
        ok = ~illegal & ~self.done
        if has_sub.any():
            keep = ok[who_s]
            if keep.any():
                sel = m & (cur > 0) & keep[:, None]                                # connected elements of the legal actions
                fl = who_s[:, None] * self.last_bus.shape[1] + psafe
                tflat[ft[sel]] = val[sel]
                lflat[fl[sel]] = val[sel]
                if reco is not None:                                              # reconnections: this end on the named busbar, the other on its last one
                    rk = reco & keep[:, None]
                    if rk.any():
                        ir = np.broadcast_to(who_s[:, None], pos.shape)[rk]
                        pr, vr = pos[rk], val[rk]
                        po = self.other_end_pos[pr]
                        self.topo[ir, pr] = vr; self.last_bus[ir, pr] = vr
                        self.topo[ir, po] = self.last_bus[ir, po]
                        aff_l[ir, self.line_of_pos[pr]] = True
                who = who_s[keep & sub_impacted[who_s]]
                aff_s[who, sub_id[who]] = True
                self._topo_dirty = True
                self._cap_dirty[who_s[keep]] = True
        if has_line.any():
            who = np.flatnonzero(has_line & ok)
            if len(who):
                l = line_id[who]
                po, pe = self.line_or_pos[l], self.line_ex_pos[l]
                off = line_status[who] < 0
                self.topo[who[off], po[off]] = -1; self.topo[who[off], pe[off]] = -1
                on = ~off
                self.topo[who[on], po[on]] = self.last_bus[who[on], po[on]]
                self.topo[who[on], pe[on]] = self.last_bus[who[on], pe[on]]
                aff_l[who, l] = True
                self._topo_dirty = True
                self._cap_dirty[who] = True
        return aff_l, aff_s

    def _isolated_shunt_busbar(self) -> np.ndarray:
        """bool [B]: busbar 1 of a substation whose shunt is out of service in the grid file carries no element any more — for the
        reference an in-service bus without a voltage: ``runpf`` fails with "Isolated bus" (pandaPowerBackend.py:1241-1244) and the
        episode ends (baseEnv.py:3523-3524)."""
        if self._iso_cols is None:            # one gather over the positions of those substations + a segmented "any"
            keep = [pos for pos, other in zip(self._shunt_off_pos, self._shunt_off_other) if not other]
            # (other: an in-service shunt / the hidden slack unit sits on that substation, never moved by this driver)
            self._iso_cols = np.concatenate(keep) if keep else np.zeros(0, dtype=np.int64)
            self._iso_seg = np.cumsum([0] + [len(p) for p in keep])[:-1]
        if len(self._iso_cols) == 0:
            return np.zeros(self.batch, dtype=bool)
        on1 = (self.topo[:, self._iso_cols] == 1).astype(np.uint8)
        return (np.add.reduceat(on1, self._iso_seg, axis=1) == 0).any(axis=1)

    def _apply_environment(self):
        """maintenance / hazards of the row every instance is about to solve: those lines are out"""
        gm = self.gm
        forced = None
        r = self.row % self.chron.shape[1]
        if self._maint is not None:
            forced = self._maint[self.scen, r]
        if self._haz is not None:
            h = self._haz[self.scen, r]
            forced = h if forced is None else (forced | h)
        if forced is None:
            return
        forced = forced & ~self.done[:, None]
        ii, ll = np.nonzero(forced)
        if len(ii):
            was_on = (self.topo[ii, self.line_or_pos[ll]] > 0) | (self.topo[ii, self.line_ex_pos[ll]] > 0)
            self.topo[ii, self.line_or_pos[ll]] = -1
            self.topo[ii, self.line_ex_pos[ll]] = -1
            if was_on.any():
                self._topo_dirty = True
                self._cap_dirty[ii[was_on]] = True

    def step(self, sub_id=None, sub_bus=None, line_id=None, line_status=None, from_reset: bool = False):
        """One env.step of every instance.  ``sub_id`` int [B] (-1: none) with ``sub_bus`` int8 [B, >= max_sub_size] (busbar per
        element of that substation in topology-vector order, 0 = leave); ``line_id`` int [B] (-1: none) with ``line_status``
        int [B] (+1 reconnect / -1 disconnect).  Returns ``(rho [B, n_line], done [B], info)``; the full records stay on the
        device (:meth:`fetch`)."""
        gm = self.gm
        aff_l, aff_s = self._apply_agent(sub_id, sub_bus, line_id, line_status) if not from_reset else (
            np.zeros((self.batch, gm.n_line), dtype=bool), np.zeros((self.batch, gm.n_sub), dtype=bool))
        self._apply_environment()
        if self._topo_dirty:
            self.engine.series_set_topo(self.topo)
            self._topo_dirty = False
            if self.tight_cap:
                self.nb_cap = self._bus_cap()
        if from_reset:
            self.engine.series_next_is_reset()
        self.step_device()
        out, status, iters, rho = self.engine.series_fetch(want_out=False)
        disc = None
        if self.protections:
            st = self.engine.series_fetch_state()
            disc = st["disc_lines"]
            tripped = (disc >= 0) & ~self.done[:, None]
            ii, ll = np.nonzero(tripped)
            if len(ii):                       # _backend_action.update_state(disc_lines): the tripped lines stay out
                self.topo[ii, self.line_or_pos[ll]] = -1
                self.topo[ii, self.line_ex_pos[ll]] = -1
                # (the engine's own mirror of the series topology already holds them: no re-upload needed)
        if len(self._shunt_off):              # (after the trips of this step: the cascade's re-solves see them too, backend.py:1466-1521)
            iso = self._isolated_shunt_busbar()
            if iso.any():
                status = np.where(iso, 2, status).astype(status.dtype)          # ST_UNSUPPLIED: the reference's "Isolated bus"
        r = self.row % self.chron.shape[1]
        # ---- cooldowns (baseEnv.py:3352-3393, :2566-2597), in the reference's order --------------------------------------
        live = ~self.done
        lc = self.line_cooldown
        lc[(lc > 0) & live[:, None]] -= 1
        if disc is not None:
            lc[(disc >= 0) & live[:, None]] = self.cd_reco
        if self._maint is not None:
            mt, md = self._maint_time[self.scen, r], self._maint_dur[self.scen, r]
            first = (mt == 0) & live[:, None]
            lc[first] = np.maximum(lc[first], md[first])
        if self._haz_dur is not None:
            hd = self._haz_dur[self.scen, r]
            lc[live] = np.maximum(lc[live], hd[live])
        if not from_reset:
            if self.cd_line > 0:
                cond = aff_l & (lc < self.cd_line)
                lc[cond] = self.cd_line
            if self.cd_sub > 0:
                sc = self.sub_cooldown
                sc[(sc > 0) & live[:, None]] -= 1
                sc[aff_s] = self.cd_sub
        newly_done = live & (status != 0)
        self.done |= newly_done
        self.row = self.row + 1
        self.n_steps += 1
        return rho, self.done.copy(), {"status": status, "iters": iters, "disc_lines": disc, "newly_done": newly_done}

    def line_angles(self, out: np.ndarray):
        """``theta_or, theta_ex`` [B, n_line] of the full result records ``out`` (:meth:`fetch`) as an OBSERVATION of the reference
        shows them: the kernels' records carry 0 for open lines, PandaPowerBackend reports the angle of the bus the end was last
        attached to (it zeroes the voltage of open lines, not the angle: pandaPowerBackend.py:1163-1187), 0 when that busbar has no
        element left.  Same rule as ``B200Backend._theta_of_open_lines``, for the whole batch."""
        from .engine import OutputView
        gm, B = self.gm, self.batch
        v = OutputView(gm, out)
        th_or, th_ex = np.array(v.theta_or, dtype=np.float32), np.array(v.theta_ex, dtype=np.float32)
        on = self.line_status()
        if on.all():
            return th_or, th_ex
        ns, nh, dt = gm.n_sub, gm.n_hidden, gm.dim_topo
        slot_theta = np.full((B, gm.n_slot + 1), np.nan, dtype=np.float64)          # (last column: dump for disconnected elements)

        def scatter(sub, pos, theta):
            bus = self.topo[:, pos].astype(np.int64)
            slot = np.where(bus > 0, np.asarray(sub, dtype=np.int64)[None, :] + (bus - 1) * ns, gm.n_slot)
            np.put_along_axis(slot_theta, slot, np.asarray(theta, dtype=np.float64), axis=1)

        scatter(gm.line_or_sub, self.line_or_pos, v.theta_or); scatter(gm.line_ex_sub, self.line_ex_pos, v.theta_ex)
        scatter(gm.gen_sub, np.asarray(gm.gen_pos, dtype=np.int64), v.unit_theta[:, nh:])
        scatter(gm.load_sub, np.asarray(gm.load_pos, dtype=np.int64), v.load_theta)
        slot_theta[:, gm.n_slot] = np.nan
        for th, sub, pos in ((th_or, gm.line_or_sub, self.line_or_pos), (th_ex, gm.line_ex_sub, self.line_ex_pos)):
            last = self.last_bus[:, pos].astype(np.int64)
            a = np.take_along_axis(slot_theta, np.asarray(sub, dtype=np.int64)[None, :] + (last - 1) * ns, axis=1)
            th[~on] = np.where(np.isfinite(a), a, 0.0)[~on]
        return th_or, th_ex

    def reset_instances(self, idx, rows=None):
        """``env.reset()`` of the instances ``idx`` (typically the finished ones): plain topology, cooldowns / counters / done flag
        cleared on the host and on the device, next chronics row ``rows`` (default: the row each instance started from).  Their
        next :meth:`step` is an ordinary step on that row (use ``rows - 1`` semantics of your agent loop accordingly)."""
        idx = np.asarray(idx, dtype=np.int64)
        if len(idx) == 0:
            return
        rows = self.t0[idx].astype(np.int64) if rows is None else np.asarray(rows, dtype=np.int64)
        self.topo[idx] = self.topo0[idx]
        self.last_bus[idx] = np.where(self.topo0[idx, :self.gm.dim_topo] > 0, self.topo0[idx, :self.gm.dim_topo], 1)
        self.line_cooldown[idx] = 0; self.sub_cooldown[idx] = 0
        self.done[idx] = False
        self.row[idx] = rows
        self.engine.series_reset_instances(idx, t_new=rows % self.chron.shape[1], topo_rows=self.topo[idx])
        self._cap_dirty[idx] = True
        if self.tight_cap:
            self.nb_cap = self._bus_cap()

    def reset_step(self):
        """the solve ``env.reset()`` performs on the first row (no soft-overflow counting, no cooldown from actions)"""
        return self.step(from_reset=True)


def random_substation_actions(env: BatchedEnv, rng: np.random.Generator):
    """BASELINE.json configs[2] / SURVEY.md 8(d) config 3: every instance draws one substation and a uniformly random busbar
    in {1, 2} for each of its elements (``MAX_SUB_CHANGED = 1``).  -> (sub_id [B], sub_bus [B, max_sub_size])"""
    B = env.batch
    sub = rng.integers(0, env.gm.n_sub, B)
    bus = rng.integers(1, 3, (B, env.max_sub_size)).astype(np.int8)
    bus[np.arange(env.max_sub_size)[None, :] >= env.sub_size[sub][:, None]] = 0
    return sub, bus
