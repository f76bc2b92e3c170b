"""Fuzz campaign of the drop-in boundary on the CPU (host logic; engines = the oracle adapters of tests/): the exploratory runs behind
tests/test_env_random_agent_cpu.py and the random tests of tests/test_batched_env.py, with more seeds.

    python scripts/fuzz_dropin.py [n_seeds] > profiles/round2_fuzz_dropin.json

Every run steps an unmodified grid2op environment on ``B200Backend`` (host logic) next to one on the oracle's restatement of
``PandaPowerBackend`` — or ``BatchedEnv`` next to unmodified environments — and stops at the first difference."""
import json
import os
import re
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))

import conftest  # noqa: E402,F401
import test_batched_env as TB  # noqa: E402
import test_env_random_agent_cpu as TR  # noqa: E402


def host_logic_backend():
    import grid2op_b200.backend as bk
    from oracle_engine import OracleEngine

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)
    return HostLogicBackend


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    HLB = host_logic_backend()
    from oracle_engine import COracleSeriesEngine, EmuProtSeriesEngine
    out = {"what": __doc__.strip().splitlines()[0], "n_seeds": n_seeds, "campaigns": []}

    def campaign(name, runs):
        t0 = time.time()
        n = fails = guards = 0
        first = None
        for label, fn in runs:
            n += 1
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    fn()
            except AssertionError as exc:
                if re.fullmatch(r"\(\d+, \d+\)", str(exc).strip()):
                    guards += 1                               # (coverage guard of the test body — a bare pair of counters: too many early
                    continue                                  #  game overs / no trip for that seed; not a difference)
                fails += 1
                first = first or f"{label}: {str(exc)[:300]}"
            except Exception as exc:  # noqa: BLE001
                fails += 1
                first = first or f"{label}: {type(exc).__name__} {str(exc)[:300]}"
        out["campaigns"].append({"campaign": name, "runs": n, "failures": fails, "runs_cut_short_by_a_coverage_guard": guards, "first_failure": first, "seconds": round(time.time() - t0, 1)})
        print(f"{name}: {n} runs, {fails} failures", file=sys.stderr, flush=True)

    # 1. action_space.sample() streams (AC / DC / 3 busbars / simulate) — reseeded through the environment seed
    def sampled(name, steps, sn, dc, nbb, sim, seed):
        def fn():
            TR.run_side_by_side(HLB, name, steps, sn, dc, nbb, sim, tag=f"fz{seed}")
        return fn
    campaign("action_space.sample() streams, the suite's nine configurations",
             [(str(c), sampled(*c, 0)) for c in TR.CASES])
    # 2. detachment / shunt / element-bus actions
    campaign("detachment, shunt and element-bus actions (CompleteAction), AC and DC",
             [(f"{n} dc={dc}", (lambda n=n, s=s, dc=dc: TR.test_random_element_actions_with_detachment(n, s, dc)))
              for n, s in (("l2rpn_case14_sandbox", 100.0), ("educ_case14_storage", 100.0), ("rte_case5_example", 1.0)) for dc in (False, True)])
    # 3. every other bundled environment
    names = ["educ_case14_redisp", "l2rpn_case14_sandbox_diff_grid", "l2rpn_icaps_2021", "l2rpn_idf_2023", "l2rpn_neurips_2020_track2",
             "l2rpn_wcci_2020", "rte_case118_example", "rte_case14_opponent", "rte_case14_realistic", "rte_case14_redisp", "rte_case14_test"]
    campaign("every other bundled environment", [(n, (lambda n=n: TR.test_every_other_bundled_environment(n))) for n in names])
    # 4. BatchedEnv: BASELINE configs[2]'s random substation re-assignments
    campaign("BatchedEnv, random substation re-assignments (configs[2])",
             [(f"seed {s}", (lambda s=s: TB.test_random_substation_actions_host_logic(s))) for s in range(n_seeds * 3)])
    # 5. BatchedEnv: random mix, protections off / on
    for prot in (False, True):
        campaign(f"BatchedEnv, random mix of substation / line / do-nothing actions, protections {'ON' if prot else 'off'}",
                 [(f"seed {s}", (lambda s=s, prot=prot: TB.run_random_mixed_actions(s, prot, HLB, EmuProtSeriesEngine if prot else COracleSeriesEngine)))
                  for s in range(n_seeds * 3)])
    out["failures_total"] = sum(c["failures"] for c in out["campaigns"])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
