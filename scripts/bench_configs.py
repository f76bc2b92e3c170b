#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs that are parity-test cases rather than the bench line
(configs 3, 4, 5) + the protections-on variant of config 2.  Prints one JSON object; run on the GPU box:
    python scripts/bench_configs.py > gpurun_out/other_configs.json
Data: grid fixtures of tests/golden (file state of each grid), injections jittered per instance
(numpy default_rng(0), loads/gens x U(0.95,1.05)); topology changes for config 3 as in SURVEY.md 8(d)."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
GOLD = os.path.join(REPO, "tests", "golden")

from grid2op_b200.engine import PowerFlowEngine  # noqa: E402
from grid2op_b200.gridmodel import GridModel  # noqa: E402
from grid2op_b200.rollout import BatchedDoNothing  # noqa: E402


def jitter(gm, n, seed=0):
    rng = np.random.default_rng(seed)
    inj = np.tile(gm.default_inj(), (n, 1))
    sl = gm.inj_slices()
    for k in ("load_p", "load_q", "gen_p"):
        inj[:, sl[k]] *= rng.uniform(0.95, 1.05, (n, inj[:, sl[k]].shape[1]))
    return inj


def synth_chron(gm, n_rows=16, seed=0):
    """Synthetic time series for grids without a bundled chronics fixture: the grid file's own set points, loads and
    generations jittered row by row (float32 [1, n_rows, 2 n_load + 2 n_gen], backend element order)."""
    sl, inj0 = gm.inj_slices(), gm.default_inj()
    row = np.concatenate([inj0[sl["load_p"]], inj0[sl["load_q"]], inj0[sl["gen_p"]], inj0[sl["gen_vm"]] * gm.prod_pu_to_kv])
    chron = np.tile(row, (1, n_rows, 1))
    n = 2 * gm.n_load + gm.n_gen
    chron[..., :n] *= np.random.default_rng(seed).uniform(0.97, 1.03, (1, n_rows, n))
    return chron.astype(np.float32)


def series_rate(gm, B, policy, steps=20):
    """Device-resident DoNothing stepping (nothing crosses PCIe): seconds per step, CUDA-synchronised wall clock."""
    os.environ["B200PF_PLAN_POLICY"] = str(policy)
    env = BatchedDoNothing(gm, synth_chron(gm), B)
    for _ in range(3):
        env.step_device()
    env.engine.sync()
    t = time.perf_counter()
    for _ in range(steps):
        env.step_device()
    env.engine.sync()
    dt = (time.perf_counter() - t) / steps
    _, status, iters, _ = env.fetch()
    info = {"seconds_per_step": dt, "env_step_per_s": B / dt, "converged_fraction": float((status == 0).mean()),
            "mean_iters": float(iters[status == 0].mean()) if (status == 0).any() else None,
            "launch": env.engine.last_launch_info(), "kernel": env.engine.plan_stats()}
    env.close()
    os.environ.pop("B200PF_PLAN_POLICY", None)
    return info


def sub_of_pos(gm):
    m = np.zeros(gm.dim_topo, dtype=np.int64)
    m[gm.line_or_pos] = gm.line_or_sub; m[gm.line_ex_pos] = gm.line_ex_sub
    m[gm.gen_pos] = gm.gen_sub; m[gm.load_pos] = gm.load_sub
    if gm.n_storage:
        m[gm.storage_pos] = gm.storage_sub
    return m


def time_run(eng, topo, inj, reps=5, **kw):
    cap = eng.max_active_buses(topo)            # host-side bound, computed once (not part of the timed call)
    t = time.perf_counter()
    eng.run(topo, inj, nb_cap=cap, **kw)
    eng.first_call_s = time.perf_counter() - t  # includes building the topology plans of the planned kernel
    eng.run(topo, inj, nb_cap=cap, **kw)
    eng.run(topo, inj, nb_cap=cap, **kw)
    t = time.perf_counter()
    for _ in range(reps):
        out, status, iters, _ = eng.run(topo, inj, nb_cap=cap, **kw)
    dt = (time.perf_counter() - t) / reps
    return dt, status, iters


def main():
    res = {}
    # ---- config 3: 36 substations, batch 1024, random one-substation topology per instance
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_neurips_2020_track1.npz"))
    B = 1024
    rng = np.random.default_rng(1)
    topo = np.tile(gm.default_topo(), (B, 1))
    sp = sub_of_pos(gm)
    for i in range(B):
        s = rng.integers(0, gm.n_sub)
        pos = np.flatnonzero(sp == s)
        topo[i, pos] = rng.integers(1, 3, len(pos))
    inj = jitter(gm, B)
    for pol, tag in ((1, "pivoting"), (2, "planned")):
        eng = PowerFlowEngine(gm, max_batch=B)
        eng.set_kernel_policy(pol)
        dt, status, iters = time_run(eng, topo, inj)
        res["config3_36sub_batch1024_random_topology_" + tag] = {
            "seconds_per_batch_host_call": dt, "env_step_per_s_e2e_host_buffers": B / dt, "converged_fraction": float((status == 0).mean()),
            "mean_iters": float(iters[status == 0].mean()), "launch": eng.last_launch_info(), "kernel": eng.plan_stats(),
            "first_call_seconds": eng.first_call_s}
        eng.close()
    for pol, tag in ((1, "pivoting"), (2, "planned")):
        res["config3grid_36sub_batch1024_donothing_device_resident_" + tag] = series_rate(gm, 1024, pol, steps=20 if pol == 1 else 100)
    if True:   # protections on a grid beyond the warp kernel: planned kernel + host re-planning of tripped instances
        env = BatchedDoNothing(gm, synth_chron(gm), 1024, protections=True)
        env.reset_step()
        env.engine.sync()
        t = time.perf_counter()
        for _ in range(50):
            env.step_device()
        env.engine.sync()
        dt = (time.perf_counter() - t) / 50
        st = env.fetch_state()
        res["config3grid_36sub_batch1024_protections_on_device_resident_planned"] = {
            "seconds_per_step": dt, "env_step_per_s": 1024 / dt, "done_fraction_after_50_steps": float(st["done"].mean()),
            "tripped_lines": int((st["disc_lines"] >= 0).sum()), "kernel": env.engine.plan_stats()}
        env.close()
    # ---- config 5: 118 substations, batch 8192 (here 2048 per call x 4), DoNothing
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_wcci_2022_dev.npz"))
    B = 2048
    topo = np.tile(gm.default_topo(), (B, 1))
    inj = jitter(gm, B)
    for pol, tag in ((1, "pivoting"), (2, "planned")):
        eng = PowerFlowEngine(gm, max_batch=B)
        eng.set_kernel_policy(pol)
        dt, status, iters = time_run(eng, topo, inj, reps=2 if pol == 1 else 10)
        res["config5_118sub_batch2048_donothing_" + tag] = {
            "seconds_per_batch_host_call": dt, "env_step_per_s_e2e_host_buffers": B / dt, "converged_fraction": float((status == 0).mean()),
            "mean_iters": float(iters[status == 0].mean()), "launch": eng.last_launch_info(), "kernel": eng.plan_stats(),
            "first_call_seconds": eng.first_call_s}
        eng.close()
    for pol, tag in ((1, "pivoting"), (2, "planned")):
        res["config5_118sub_batch2048_donothing_device_resident_" + tag] = series_rate(gm, 2048, pol, steps=5 if pol == 1 else 100)
    for var, what in ((128, "T128"), (64, "T64"), (32, "T32")):
        os.environ["B200PF_SPARSE_T"] = str(var)
        res["config5_118sub_batch2048_device_resident_planned_" + what] = series_rate(gm, 2048, 2, steps=50)
        res["config5_118sub_batch8192_device_resident_planned_" + what] = series_rate(gm, 8192, 2, steps=20)
    os.environ.pop("B200PF_SPARSE_T", None)
    # ---- config 4: case14 N-1 sweep, 4096 base states x 20 outages
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    B = 4096
    topo = np.tile(gm.default_topo(), (B, 1))
    inj = jitter(gm, B)
    for pol, tag in ((1, "pivoting"), (2, "planned")):
        eng = PowerFlowEngine(gm, max_batch=B * gm.n_line)
        eng.set_kernel_policy(pol)
        eng.n1_sweep(topo, inj)
        t = time.perf_counter()
        reps = 3
        for _ in range(reps):
            rho, status = eng.n1_sweep(topo, inj)
        dt = (time.perf_counter() - t) / reps
        res["config4_case14_n1_sweep_4096x20_" + tag] = {
            "seconds_per_sweep_host_call": dt, "contingencies_per_s": B * gm.n_line / dt, "converged_fraction": float((status == 0).mean()),
            "launch": eng.last_launch_info(), "kernel": eng.plan_stats()}
        eng.close()
    # ---- config 2 with protections on (device-side cascading failure), device-resident
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    th = np.array([541.0, 450.0, 375.0, 636.0, 175.0, 285.0, 335.0, 657.0, 496.0, 827.0, 442.0, 641.0, 840.0, 156.0, 664.0, 235.0,
                   119.0, 179.0, 1986.0, 1572.0], dtype=np.float32)       # thermal limits of the sandbox config (config.py:17-39)
    env = BatchedDoNothing(gm, chron, 4096, protections=True, thermal_limit_a=th)
    env.reset_step()
    env.engine.sync()
    t = time.perf_counter()
    steps = 200
    for _ in range(steps):
        env.step_device()
    env.engine.sync()
    dt = (time.perf_counter() - t) / steps
    st = env.fetch_state()
    res["config2_case14_batch4096_protections_on"] = {"seconds_per_step": dt, "env_step_per_s": 4096 / dt,
                                                      "done_fraction_after_200_steps": float(st["done"].mean())}
    env.close()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
