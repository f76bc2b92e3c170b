"""Batched DoNothing stepping over chronics: the throughput driver next to the drop-in Backend.

What one ``step`` restates for a whole batch of independent environments (reference, per env):
``BaseEnv.step`` grid2op/Environment/baseEnv.py:3778-3872 restricted to the DoNothing agent with
``NO_OVERFLOW_DISCONNECTION`` (the setting of the reference's own profiling script,
``_profiling/profiler_do_nothing.py:52``): next chronics row (gridStateFromFile.py:766-776) ->
``apply_action`` -> ``runpf`` -> read-back + ``rho = a_or / thermal_limit``
(``Backend.get_relative_flow`` backend.py:1145-1168).  An instance whose power flow fails is "done".

Two ways to drive it:
* :meth:`step_device` - chronics resident in HBM, nothing crosses PCIe (``value`` in bench.py);
* :meth:`step_host`   - host buffers in, host buffers out through the C ABI (``e2e`` in bench.py).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import PowerFlowEngine
from .gridmodel import GridModel

__all__ = ["BatchedDoNothing", "instance_schedule"]


def instance_schedule(n: int, n_scen: int, n_rows: int, offset: int = 0):
    """Deterministic, seed-free assignment used by the benchmarks:
    instance i replays scenario ``i mod n_scen`` starting at row ``(i*37) mod n_rows``."""
    i = np.arange(offset, offset + n, dtype=np.int64)
    return (i % n_scen).astype(np.int32), ((i * 37) % n_rows).astype(np.int32)


class BatchedDoNothing:
    def __init__(self, gm: GridModel, chron: np.ndarray, batch: int, device: int = 0, offset: int = 0,
                 max_iter: int = 10, tol_mva: float = 1e-8, is_dc: bool = False, scen: Optional[np.ndarray] = None,
                 t0: Optional[np.ndarray] = None, thermal_limit_a: Optional[np.ndarray] = None,
                 protections: bool = False, hard_overflow_threshold: float = 2.0, soft_overflow_threshold: float = 1.0,
                 nb_timestep_overflow_allowed: int = 2, engine=None):
        self.gm = gm
        self.batch = int(batch)
        self.chron = np.ascontiguousarray(chron, dtype=np.float32)
        assert self.chron.ndim == 3 and self.chron.shape[2] == 2 * gm.n_load + 2 * gm.n_gen
        self.scen, self.t0 = instance_schedule(self.batch, self.chron.shape[0], self.chron.shape[1], offset)
        if scen is not None:
            self.scen = np.ascontiguousarray(scen, dtype=np.int32)
        if t0 is not None:
            self.t0 = np.ascontiguousarray(t0, dtype=np.int32)
        self.thermal_limit_a = np.ascontiguousarray(gm.thermal_limit_a if thermal_limit_a is None else thermal_limit_a,
                                                    dtype=np.float32)
        # (engine: an object with the series API of PowerFlowEngine — tests inject the oracle adapter to check host logic on CPU)
        self.engine = engine if engine is not None else PowerFlowEngine(gm, max_batch=self.batch, device=device)
        self.max_iter, self.tol_mva, self.is_dc = int(max_iter), float(tol_mva), bool(is_dc)
        self.topo0 = np.tile(gm.default_topo(), (self.batch, 1))
        self.nb_cap = int(np.count_nonzero(np.bincount(self._slots(gm.default_topo()), minlength=gm.n_slot)))
        self.engine.series_bind(self.chron, self.scen, self.t0, gm.default_inj(), self.thermal_limit_a)
        self.engine.series_set_topo(self.topo0)
        self.protections = bool(protections)
        if self.protections:
            # Backend.next_grid_state on the device: hard / soft overflow disconnections + cascade re-solves
            self.engine.series_protections(True, hard_overflow_threshold, soft_overflow_threshold, nb_timestep_overflow_allowed)
        # host path state
        self._t_host = self.t0.astype(np.int64).copy()
        self._stage = None
        self._collated = None
        self._k = 0
        self.host_chunks = 2          # measured on B200 + PCIe Gen5: 2 chunks 222 us, 1: 271 us, 4: 227 us, 8: 296 us per 4096-step
        self._sl = gm.inj_slices()
        self._inj0 = gm.default_inj()

    def _slots(self, tv):
        gm = self.gm
        subs = np.concatenate([gm.line_or_sub, gm.line_ex_sub, gm.gen_sub, gm.load_sub, gm.storage_sub, gm.shunt_sub, gm.hidden_sub])
        pos = np.concatenate([gm.line_or_pos, gm.line_ex_pos, gm.gen_pos, gm.load_pos, gm.storage_pos,
                              gm.dim_topo + np.arange(gm.n_shunt), gm.dim_topo + gm.n_shunt + np.arange(gm.n_hidden)])
        b = tv[pos].astype(np.int64)
        ok = b > 0
        return subs[ok].astype(np.int64) + (b[ok] - 1) * gm.n_sub

    # ---- device-resident ------------------------------------------------------------------------
    def step_device(self) -> None:
        """Asynchronous: one fused kernel for the whole batch; results stay in HBM."""
        self.engine.series_step(is_dc=self.is_dc, max_iter=self.max_iter, tol_mva=self.tol_mva, nb_cap=self.nb_cap)
        self._n_dev_steps = getattr(self, "_n_dev_steps", 0) + 1

    def simulate_forecast(self, forecast: np.ndarray, forecast_has_prod_v: bool = True, topo: Optional[np.ndarray] = None):
        """Batched ``obs.simulate(do_nothing, time_step=1)`` (reference grid2op/Observation/baseObservation.py:3365-3669 fed by
        ``GridStateFromFileWithForecasts.forecasts``, gridStateFromFileWithForecasts.py:311-353): ONE launch solves, for every
        instance, the forecast row that belongs to the step it has just done (``forecast[scen, current row]``, float32
        [n_scen, n_rows, 2 n_load + 2 n_gen] from :func:`grid2op_b200.chronics.load_forecasts`) on its current topology
        (``topo``: int8 [batch, n_topo_in] for a what-if on another topology).  A forecast folder without ``prod_v_forecasted``
        keeps the current voltage set points, like the reference.  Does not advance the series; the device-side result buffers of
        the last step are overwritten (fetch them first).  Returns ``(out, status, rho)`` on the host."""
        gm, B = self.gm, self.batch
        n_rows = self.chron.shape[1]
        cur = (self.t0.astype(np.int64) + getattr(self, "_n_dev_steps", 0) - 1) % n_rows
        rows = np.ascontiguousarray(forecast[self.scen, cur], dtype=np.float32)
        if not forecast_has_prod_v:
            rows[:, 2 * gm.n_load + gm.n_gen:] = self.chron[self.scen, cur][:, 2 * gm.n_load + gm.n_gen:]
        st = self.engine.staging()
        st["topo"][:B] = self.topo0 if topo is None else np.ascontiguousarray(topo, dtype=np.int8).reshape(B, gm.n_topo_in)
        self.engine.rows_staging()[:B] = rows
        self.engine.set_static_inj(self._inj0)
        self.engine.run_rows_staged(B, is_dc=self.is_dc, max_iter=self.max_iter, tol_mva=self.tol_mva, nb_cap=0 if topo is not None else self.nb_cap)
        out, status = st["out"][:B].copy(), st["status"][:B].copy()
        rho = self.engine.view(out).a_or / self.thermal_limit_a[None, :]
        return out, status, rho

    def reset_step(self) -> None:
        """The step an environment performs inside ``reset()``: same solve, but no soft-overflow counting
        (reference grid2op/Backend/backend.py:1488-1490, ``_called_from_reset``)."""
        self.engine.series_next_is_reset()
        self.step_device()

    def fetch(self):
        return self.engine.series_fetch()

    def fetch_state(self):
        """protection_counter / timestep_overflow / disc_lines [batch, n_line], done [batch]."""
        return self.engine.series_fetch_state()

    # ---- host buffers in / out --------------------------------------------------------------------
    def step_host(self):
        """Next chronics row of every instance gathered on the host straight into the pinned float32 rows
        buffer, H2D (topology record + rows), kernel, D2H; returns (out view, status view) living in the
        pinned staging buffers."""
        if self._stage is None:
            self._stage = self.engine.staging()
            self._rows = self.engine.rows_staging()
            self._stage["topo"][:self.batch] = self.topo0
            self.engine.set_static_inj(self._inj0)
            n_rows = self.chron.shape[1]
            self._chron_flat = self.chron.reshape(-1, self.chron.shape[2])
            self._row_base = self.scen.astype(np.int64) * n_rows
        st = self._stage
        nch = self.host_chunks if self.batch >= 4 * self.host_chunks else 1
        if self._collated is not None:
            # pre-collated time series: the rows of this step are already contiguous in pinned memory
            rows_k = self._collated[self._k % self._collated.shape[0]]
            lo = 0
            for c in range(nch):
                hi = self.batch * (c + 1) // nch
                self.engine.rows_chunk_launch_from(lo, hi - lo, rows_k[lo:hi], is_dc=self.is_dc, max_iter=self.max_iter,
                                                   tol_mva=self.tol_mva, nb_cap=self.nb_cap)
                lo = hi
            self._k += 1
            self._t_host = (self._t_host + 1) % self.chron.shape[1]
            self.engine.rows_chunk_wait()
            return st["out"][:self.batch], st["status"][:self.batch]
        idx = self._row_base + self._t_host
        lo = 0
        for c in range(nch):                      # gather chunk c on the host while the device works on chunk c-1
            hi = self.batch * (c + 1) // nch
            np.take(self._chron_flat, idx[lo:hi], axis=0, out=self._rows[lo:hi])
            self.engine.rows_chunk_launch(lo, hi - lo, is_dc=self.is_dc, max_iter=self.max_iter, tol_mva=self.tol_mva,
                                          nb_cap=self.nb_cap)
            lo = hi
        self._t_host = (self._t_host + 1) % self.chron.shape[1]
        self.engine.rows_chunk_wait()
        return st["out"][:self.batch], st["status"][:self.batch]

    # ---- host buffers in / out, group-pipelined (asynchronous vectorised environments) -----------------------------
    def host_groups(self, n_groups: int = 4, direct_out=False):
        """Cut the batch into ``n_groups`` contiguous groups for :meth:`group_launch` / :meth:`group_wait`: while the
        caller consumes the results of one group the others are in flight (their PCIe copies overlap the kernels of
        the rest).  Every instance still sees its own results before its next step is launched."""
        if self._stage is None:
            self.step_host()                                   # initialise staging / static inputs
            self._t_host = self.t0.astype(np.int64).copy()
            self._k = 0
        n_groups = max(1, min(int(n_groups), self.batch, 16))
        self.engine.rows_group_config(direct_out=direct_out)
        self._groups = [(self.batch * g // n_groups, self.batch * (g + 1) // n_groups) for g in range(n_groups)]
        self._gk = [self._k] * n_groups
        # per-group constants of the C call (the per-step host cost bounds how many groups pay off)
        eng = self.engine
        self._g_fn = (eng.lib.b200pf_rows_group_launch, eng.lib.b200pf_rows_group_wait, eng.h)
        self._g_tail = (int(self.is_dc), int(self.max_iter), float(self.tol_mva), int(self.nb_cap))
        self._g_views = [(self._stage["out"][lo:hi], self._stage["status"][lo:hi]) for lo, hi in self._groups]
        if self._collated is not None:
            ncol = self.chron.shape[2]
            self._g_base = [self._collated.ctypes.data + 4 * ncol * lo for lo, _ in self._groups]
            self._g_stride = 4 * ncol * self.batch
        return list(self._groups)

    def group_launch(self, g: int) -> None:
        """Asynchronous: next step of group ``g`` (H2D of its topology records + chronics rows, kernel, D2H)."""
        lo, hi = self._groups[g]
        launch, _, h = self._g_fn
        if self._collated is not None:
            k = self._gk[g]
            self._gk[g] = k + 1
            rows = self._g_base[g] + self._g_stride * (k % self._collated.shape[0])
        else:
            np.take(self._chron_flat, self._row_base[lo:hi] + self._t_host[lo:hi], axis=0, out=self._rows[lo:hi])
            self._t_host[lo:hi] = (self._t_host[lo:hi] + 1) % self.chron.shape[1]
            rows = None
        rc = launch(h, g, lo, hi - lo, rows, *self._g_tail)
        if rc:
            self.engine._check(rc, "b200pf_rows_group_launch")

    def group_wait(self, g: int):
        """Blocks until the results of group ``g`` are in the pinned staging buffers; returns (out, status) views."""
        rc = self._g_fn[1](self._g_fn[2], g)
        if rc:
            self.engine._check(rc, "b200pf_rows_group_wait")
        return self._g_views[g]

    def precollate(self, max_bytes: int = 2 << 30) -> bool:
        """Data-pipeline step done once: lay the time series out step-major in pinned host memory,
        ``[n_rows, batch, ncol]`` (row k = what every instance needs at its k-th step), so that a step ships its
        inputs with plain asynchronous copies and no host-side gather.  Returns False (and changes nothing) when the
        buffer would exceed ``max_bytes``."""
        n_rows, ncol = self.chron.shape[1], self.chron.shape[2]
        if n_rows * self.batch * ncol * 4 > max_bytes:
            return False
        if self._stage is None:
            self.step_host()                                   # initialise staging / static inputs
            self._t_host = self.t0.astype(np.int64).copy()
        buf = self.engine.pinned_empty((n_rows, self.batch, ncol), np.float32)
        flat = self.chron.reshape(-1, ncol)
        base = self.scen.astype(np.int64) * n_rows
        for k in range(n_rows):
            np.take(flat, base + (self._t_host + k) % n_rows, axis=0, out=buf[k])
        self._collated = buf
        self._k = 0
        return True

    def bytes_per_step_host(self):
        gm = self.gm
        return self.batch * (gm.n_topo_in + 4 * (2 * gm.n_load + 2 * gm.n_gen)), self.batch * (4 * gm.n_out + 8)

    def close(self):
        self.engine.close()
