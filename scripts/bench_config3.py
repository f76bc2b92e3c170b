#!/usr/bin/env python
"""BASELINE.json configs[2], measured as SURVEY.md 8(d) states it: l2rpn_neurips_2020_track1 (36 substations) AC, batch 1024,
every step EVERY instance draws one substation and a uniformly random busbar assignment for its elements
(numpy default_rng(1)), applied on top of its current topology (the topologies random-walk: after a few steps every instance
sits in a topology nobody has seen before), then apply -> solve -> read back (rho + done flags on the host), opponent and
protections off, chronics = the bundled scenarios.  Instances whose power flow fails are game over and restart from the
plain topology (counted).  Reported: env.step/s of the first step (cold plan cache), of the steady state, the plan-cache hit
rate, plans built per step, host threads, convergence rate — for the planned kernel with host-side planning (policy 2) and
for the pivoting kernels that discover the topology on the device (policy 1).

    python scripts/bench_config3.py [--batch 1024] [--steps 40] > gpurun_out/config3.json"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from grid2op_b200.batched_env import BatchedEnv, random_substation_actions  # noqa: E402
from grid2op_b200.gridmodel import GridModel  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def run(policy, batch, steps, grid, chronf, seed=1, tight_cap=True):
    gm = GridModel.from_npz(os.path.join(GOLD, grid))
    chron = np.load(os.path.join(GOLD, chronf))["chron"]
    os.environ["B200PF_PLAN_POLICY"] = str(policy)
    env = BatchedEnv(gm, chron, batch, tight_cap=tight_cap)
    env.engine.set_kernel_policy(policy)
    rng = np.random.default_rng(seed)
    env.reset_step()
    ts, conv, restarts, caps = [], [], 0, []
    c0 = env.engine.plan_counters()
    per_step = []
    for k in range(steps):
        sub, bus = random_substation_actions(env, rng)
        t = time.perf_counter()
        rho, done, info = env.step(sub, bus)
        ts.append(time.perf_counter() - t)
        caps.append(int(env.nb_cap))
        conv.append(float((info["status"] == 0).mean()))
        c1 = env.engine.plan_counters()
        per_step.append({k2: c1[k2] - c0[k2] for k2 in c1}); c0 = c1
        if done.any():                                   # game over -> that environment restarts from the plain topology
            who = np.flatnonzero(done)
            restarts += len(who)
            env.reset_instances(who, rows=env.row[who])      # fresh episode on the plain topology, chronics go on
    ts = np.array(ts)
    steady = ts[len(ts) // 2:]
    look = sum(p["lookups"] for p in per_step[len(ts) // 2:]); hits = sum(p["hits"] for p in per_step[len(ts) // 2:])
    res = {"policy": {0: "automatic (planned kernel unless more than 1/8 of the call's instances need a new plan)",
                      1: "pivoting kernels (topology discovered on the device)", 2: "planned kernel (topology plans built on the host)"}[policy],
           "batch": batch, "steps": steps, "kernel": env.engine.plan_stats(), "launch": env.engine.last_launch_info(),
           "env_step_per_s_first_step": batch / ts[0], "env_step_per_s_steady": batch / float(np.median(steady)),
           "ms_per_step_steady_median": 1e3 * float(np.median(steady)), "ms_per_step_min": 1e3 * float(ts.min()),
           "plan_cache_hit_rate_steady": (hits / look) if look else None,
           "plans_built_per_step_steady": float(np.mean([p["built"] for p in per_step[len(ts) // 2:]])),
           "cache_resets": env.engine.plan_counters()["cache_resets"],
           "converged_fraction_mean": float(np.mean(conv)), "game_over_restarts": restarts,
           "bus_cap": {"tight": bool(tight_cap), "min": min(caps), "max": max(caps), "every_slot": int(gm.n_slot)},
           "host_threads_visible": os.cpu_count(), "plan_threads_cap": os.environ.get("B200PF_PLAN_THREADS", "min(64, hardware)")}
    env.close()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--grid", default="gridmodel_l2rpn_neurips_2020_track1.npz")
    ap.add_argument("--chron", default="neurips_2020_track1_chronics.npz")
    a = ap.parse_args()
    out = {"workload": "l2rpn_neurips_2020_track1 (36 substations) AC, batch %d, one random substation re-assignment per instance and step, "
                       "host buffers for the actions, rho + status read back every step" % a.batch}
    for pol in (2, 1, 0):
        out[f"policy{pol}"] = run(pol, a.batch, a.steps, a.grid, a.chron)
        print(pol, json.dumps(out[f"policy{pol}"]), file=sys.stderr, flush=True)
    # the pivoting kernels with their workspace sized for every bus slot (round-2 state before the bus cap followed the topology)
    out["policy1_every_slot"] = run(1, a.batch, a.steps, a.grid, a.chron, tight_cap=False)
    print("1 (every slot)", json.dumps(out["policy1_every_slot"]), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))
