#!/usr/bin/env python
"""bench.py - env.step()/s of the batched power-flow hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle's C restatement on the host cores

Workload at N=1 (BASELINE.json configs[1]): l2rpn_case14_sandbox, AC Newton-Raphson, batch 4096 per
GPU, DoNothing rollout over the bundled chronics (instance i -> scenario i mod 3, start row (i*37) mod
576), NO_OVERFLOW_DISCONNECTION like the reference's own profiling script.  A "step" is
one pass of the hot path over the whole batch = batch env.step() calls.  N>1: weak scaling (4096 per
GPU), no data-path collective inside the solve, one NCCL gather of rho per step to rank 0
(asynchronous, overlapped with the next step's solve).

One JSON line on stdout (rank 0).  Keys documented in DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLD = os.path.join(REPO, "tests", "golden")

METRIC = "env.step()/sec at batch on l2rpn_case14 AC"
UNIT = "env.step()/s"


def load_workload():
    from grid2op_b200.gridmodel import GridModel
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    return gm, chron


def flops_view(gm, mean_iters, batch, kern_s, clocks):
    """SURVEY.md 8(d), second view: dense-LU flops per instance-iteration (2/3 d^3 + 2 d^2, d = NR dimension with all
    elements on busbar 1) against the fp32 CUDA-core peak at the SM clock sampled during the run (148 SMs x 128 FMA
    lanes x 2; the solve is 22x22 per instance, not tensor-core shaped, so MEASURED_PEAKS' bf16 figure does not apply)."""
    n_ref = 1
    n_pv = int(len(set(int(x) for x in gm.gen_sub))) - n_ref
    d = n_pv + 2 * (gm.n_sub - n_pv - n_ref)
    fl = (2.0 / 3.0) * d ** 3 + 2.0 * d ** 2
    ach = batch * fl * mean_iters / kern_s / 1e12
    mhz = (clocks or {}).get("sm_mhz") or 1965
    peak = 148 * 128 * 2 * mhz * 1e6 / 1e12
    return {"nr_dimension": d, "flops_per_instance_iteration": fl, "achieved_tflops": ach, "peak_fp32_cuda_core_tflops": peak,
            "frac": ach / peak}


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(gm, mean_iters):
    """B_iter (state round-tripped through HBM once per NR iteration, int8 topo, fp64
    V as we keep it) x iterations + result write-back.  Also the compulsory bytes of THIS design (fused:
    inputs read once, results written once)."""
    n_bus_max = gm.n_slot
    inj_vals = 2 * gm.n_load + 2 * gm.n_gen + gm.n_storage + 3 * gm.n_shunt
    b_iter = gm.dim_topo + 4 * inj_vals + 2 * 16 * n_bus_max + 4
    result = 4 * gm.n_out
    survey = b_iter * mean_iters + result
    compulsory = 4 * (2 * gm.n_load + 2 * gm.n_gen) + gm.n_topo_in + 8 + 4 * gm.n_out + 4 * gm.n_line + 8
    return float(survey), float(compulsory), int(b_iter)


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(gm, chron, n_inst, seconds_budget=12.0, nthreads=0):
    """Times oracle/pf_oracle.c (OpenMP over instances) on chronics rows of the same workload."""
    from oracle.c_oracle import COracle
    from grid2op_b200.rollout import instance_schedule
    orc = COracle(gm, nthreads=nthreads)
    scen, t0 = instance_schedule(n_inst, chron.shape[0], chron.shape[1])
    sl = gm.inj_slices()
    nl, ng = gm.n_load, gm.n_gen
    topo = np.tile(gm.default_topo(), (n_inst, 1))
    inj = np.tile(gm.default_inj(), (n_inst, 1))
    done, t_used, k = 0, 0.0, 0
    orc.run(topo[:256], inj[:256])   # warm-up (thread pool, page faults)
    while t_used < seconds_budget:
        rows = chron[scen, (t0 + k) % chron.shape[1]]
        inj[:, sl["load_p"]] = rows[:, :nl]; inj[:, sl["load_q"]] = rows[:, nl:2 * nl]
        inj[:, sl["gen_p"]] = rows[:, 2 * nl:2 * nl + ng]
        inj[:, sl["gen_vm"]] = rows[:, 2 * nl + ng:] / gm.prod_pu_to_kv[None, :]
        t = time.perf_counter()
        out, status, iters, _ = orc.run(topo, inj)
        t_used += time.perf_counter() - t
        assert (status == 0).all()
        done += n_inst; k += 1
    return done / t_used, orc.max_threads if nthreads <= 0 else nthreads, done, t_used


def run_reference(args):
    """CPU arm.  The reference's implementation of this path is pandapower (pure Python, third party,
    not installable here - no wheel, no network); the arm therefore times the oracle's C restatement of
    the same algorithm on all host cores: a (much) faster stand-in than the original, labelled 'port'."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gm, chron = load_workload()
    batch = args.batch
    from oracle.c_oracle import COracle
    from grid2op_b200.rollout import instance_schedule
    scen, t0 = instance_schedule(batch, chron.shape[0], chron.shape[1])
    sl = gm.inj_slices(); nl, ng = gm.n_load, gm.n_gen
    topo = np.tile(gm.default_topo(), (batch, 1)); inj = np.tile(gm.default_inj(), (batch, 1))
    # "all the host threads it can use": try the logical and the physical core count, keep the faster
    ncpu = os.cpu_count() or 1
    best = None
    for nt in sorted({ncpu, max(1, ncpu // 2)}, reverse=True):
        o = COracle(gm, nthreads=nt)
        o.run(topo, inj)
        t = time.perf_counter(); o.run(topo, inj); dt_ = time.perf_counter() - t
        if best is None or dt_ < best[0]:
            best = (dt_, nt, o)
    orc = best[2]
    n_threads = best[1]

    def step(k):
        rows = chron[scen, (t0 + k) % chron.shape[1]]
        inj[:, sl["load_p"]] = rows[:, :nl]; inj[:, sl["load_q"]] = rows[:, nl:2 * nl]
        inj[:, sl["gen_p"]] = rows[:, 2 * nl:2 * nl + ng]
        inj[:, sl["gen_vm"]] = rows[:, 2 * nl + ng:] / gm.prod_pu_to_kv[None, :]
        out, status, iters, _ = orc.run(topo, inj)
        return status

    for k in range(args.warmup):
        step(k)
    t = time.perf_counter()
    for k in range(args.steps):
        st = step(args.warmup + k)
    dt = time.perf_counter() - t
    value = batch * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (bundled l2rpn_case14_sandbox chronics rows replayed, DoNothing)",
        "config": {"workload": f"l2rpn_case14_sandbox AC Newton-Raphson, batch {batch}, DoNothing rollout, host CPU",
                   "batch_per_step": batch},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": "port",
                         "sample": f"{args.steps} steps x {batch} instances (whole batch per step), oracle/pf_oracle.c, OpenMP x{n_threads}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from grid2op_b200.rollout import BatchedDoNothing
    os.environ["B200PF_PLAN_POLICY"] = str(args.policy)
    gm, chron = load_workload()
    batch = args.batch
    env = BatchedDoNothing(gm, chron, batch, device=local, offset=rank * batch)
    eng = env.engine
    stream = torch.cuda.Stream()            # a real (non-default) stream shared by torch events and the engine
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    # results land in torch tensors (so NCCL can gather them); rho is double-buffered so that the gather of
    # step k (NCCL stream) overlaps the solve of step k+1 (compute stream)
    rho_bufs = [torch.empty((batch, gm.n_line), dtype=torch.float32, device="cuda") for _ in range(2)]
    status = torch.empty((batch,), dtype=torch.int32, device="cuda")
    iters = torch.empty((batch,), dtype=torch.int32, device="cuda")
    gather_lists = [[torch.empty_like(rho_bufs[0]) for _ in range(world)] for _ in range(2)] if (world > 1 and rank == 0) else [None, None]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    state = {"k": 0, "pending": None}

    def one_step():
        k = state["k"]
        buf = rho_bufs[k % 2]
        eng.series_bind_outputs(0, status.data_ptr(), iters.data_ptr(), buf.data_ptr())
        env.step_device()
        if world > 1:
            if state["pending"] is not None:
                state["pending"].wait()                 # gather of step k-1 overlapped this step's kernel
            state["pending"] = dist.gather(buf, gather_lists[k % 2], dst=0, async_op=True)
        state["k"] = k + 1

    def drain():
        if state["pending"] is not None:
            state["pending"].wait()
            state["pending"] = None

    for _ in range(max(args.warmup, 3)):
        one_step()
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = eng.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    iter_sum = 0.0
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()                       # evict L2 between timed iterations (not timed)
        a.record()
        one_step()
        b.record()
    tail_a, tail_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tail_a.record()
    drain()                                 # the last gather is timed too
    tail_b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count - launches0
    dev_ms = float(sum(a.elapsed_time(b) for a, b in evs)) + float(tail_a.elapsed_time(tail_b))
    n_bad = int((status != 0).sum().item())
    mean_iters = float(iters.float().mean().item())
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = world * batch * args.steps / (dev_ms_max * 1e-3)

    # ---- end to end through host buffers (C-ABI staged call), every rank, max over ranks ----------
    eng.set_stream(0)
    collated = env.precollate()                  # one-off data-pipeline step (untimed): step-major pinned copy of the series
    for _ in range(3):
        env.step_host()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rho_max = 0.0
    for _ in range(args.steps):
        out, st = env.step_host()
        rho_max += float((st != 0).sum())          # consume the step's result on the host (done flags; the full
        #                                            observation record `out` sits in pinned host memory)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sync = world * batch * args.steps / float(t.item())
    e2e_value, e2e_mode = e2e_sync, "lockstep"
    if args.e2e_groups >= 1:
        # asynchronous vectorised environments: the batch is cut into G groups, each with its own stream; while the host
        # consumes the results of one group (and issues its next step) the others are in flight.  Every instance still
        # gets its results back on the host before its next step is launched; K steps of every instance are timed.
        groups = env.host_groups(args.e2e_groups, direct_out=args.e2e_direct)
        ng = len(groups)
        for g in range(ng):
            env.group_launch(g)
        for _ in range(3):
            for g in range(ng):
                env.group_wait(g); env.group_launch(g)
        for g in range(ng):
            env.group_wait(g)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for g in range(ng):
            env.group_launch(g)
        for k in range(args.steps):
            for g in range(ng):
                out, st = env.group_wait(g)
                rho_max += float((st != 0).sum())
                if k + 1 < args.steps:
                    env.group_launch(g)
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * batch * args.steps / float(t.item())
        fl = args.e2e_direct
        e2e_mode = f"{ng} groups in flight" + (", results stored by the kernel straight into pinned host memory" if fl & 1 else "") \
            + (", chronics rows read by the kernel straight from pinned host memory (no copy-engine H2D)" if fl & 2 else "") \
            + (", status/iteration counts stored by the kernel into pinned host memory" if fl & 4 else "")
    h2d, d2h = env.bytes_per_step_host()
    if eng.plan_stats()["last_kernel"] == "planned_sparse":
        h2d -= batch * gm.n_topo_in            # the planned kernel works from the cached topology plan: no topology records cross PCIe

    if rank == 0:
        peak, peak_src = measured_peaks()
        surv, comp, b_iter = algorithmic_bytes(gm, mean_iters)
        kern_s = dev_ms * 1e-3 / args.steps                     # one launch per step (N=1: the event pair brackets only it)
        achieved = batch * surv / kern_s / 1e9
        traffic = None
        summ = os.path.join(REPO, "profiles", "round1_ncu_planned_case14_summary.json")
        if os.path.exists(summ):
            try:
                traffic = json.load(open(summ)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu_rate, cores, cpu_n, cpu_t = (None, None, 0, 0.0)
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu_rate, cores, cpu_n, cpu_t = cpu_port_rate(gm, chron, 8192, seconds_budget=args.cpu_seconds)
            cpu = {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{cpu_n} instance-steps of the same workload in {cpu_t:.1f} s, oracle/pf_oracle.c (fp64 dense Newton), OpenMP"}
        info = eng.last_launch_info()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (bundled l2rpn_case14_sandbox chronics rows replayed, DoNothing)",
            "config": {"workload": f"l2rpn_case14_sandbox AC Newton-Raphson, batch {batch} envs per GPU, DoNothing rollout",
                       "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"dp{world} (independent instances)",
                       "l2": "flushed between timed steps (256 MiB memset, outside the event pair)",
                       "precision": "fp64 state/mismatch/flows, fp32 Jacobian+LU, tol 1e-8 MVA, max_iter 10",
                       "launch": info, "kernel": eng.plan_stats(), "mean_newton_iterations": mean_iters, "diverged": n_bad,
                       "wall_s_incl_flush": t_wall},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "mode": e2e_mode, "lockstep_value": e2e_sync,
                    "what": "BatchedDoNothing host path (group_launch/group_wait; lockstep_value = step_host()): per step H2D of the topology records + chronics rows (from pinned host memory, "
                            + ("pre-collated step-major" if collated else "gathered on the host per step")
                            + "), kernel, D2H of the full result records, host reads the done flags of every instance before "
                              "launching its next step; `value`: the batch is stepped as staggered groups (asynchronous "
                              "vectorised envs), `lockstep_value`: all instances wait for each other every step (2 pipelined chunks)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": surv, "b_iter_bytes": b_iter,
                         "compulsory_bytes_per_env_step": comp,
                         "achieved_compulsory_gbs": batch * comp / kern_s / 1e9,
                         "flops_view": flops_view(gm, mean_iters, batch, kern_s, clocks),
                         "note": "fused whole-solve kernel: state never round-trips HBM, the kernel is issue/latency bound (see DESIGN.md)"},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--e2e-direct", type=int, default=6,
                    help="group flags (include/b200pf.h): 1 kernels store results straight into pinned host memory, 2 kernels read "
                         "the chronics rows straight from pinned host memory, 4 status / iteration counts stored straight into pinned host memory")
    ap.add_argument("--e2e-groups", type=int, default=4, help="groups in flight for the end-to-end host path (<=1: lockstep only)")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--policy", type=int, default=0, choices=[0, 1, 2],
                    help="kernel policy (include/b200pf.h): 0 auto = planned sparse kernel, 1 pivoting kernels only, 2 planned always")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
