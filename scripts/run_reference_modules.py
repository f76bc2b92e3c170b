"""Runs EVERY test module of the reference (grid2op/tests/test_*.py, unmodified, from the reference tree) with B200Backend standing where
the module expects PandaPowerBackend (tests/ref_modules_runner.py, host logic over the oracle adapter, one process per module) and
rewrites profiles/round2_reference_test_modules.json.  Reasons ("why") of modules that are not green are carried over from the
previous file when the module is still not green.

    python scripts/run_reference_modules.py [n_parallel]
"""
import concurrent.futures as cf
import glob
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_TESTS = "/root/reference/grid2op/tests"
OUT = os.path.join(REPO, "profiles", "round2_reference_test_modules.json")


def run_one(mod):
    try:
        r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "ref_modules_runner.py"), "hostlogic", mod], cwd=os.path.join(REPO, "tests"),
                           capture_output=True, text=True, timeout=1800)
    except subprocess.TimeoutExpired:
        return {"module": mod, "import_error": "timeout (1800 s)"}
    for ln in r.stdout.splitlines():
        if ln.startswith("RESULT "):
            return json.loads(ln[7:])
    return {"module": mod, "import_error": "no result: " + r.stderr[-300:]}


def main():
    n_par = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    mods = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(REF_TESTS, "test_*.py")))
    old = json.load(open(OUT)) if os.path.exists(OUT) else {}
    why = {m["module"]: m.get("why", "") for m in old.get("modules_with_failures", [])}
    with cf.ThreadPoolExecutor(n_par) as ex:
        res = list(ex.map(run_one, mods))
    green, bad, noimp = [], [], []
    n_run = n_bad = n_skip = 0
    for d in res:
        if "import_error" in d:
            noimp.append({"module": d["module"], "error": d["import_error"][:300]})
            continue
        n_run += d["run"]; n_bad += d["fail"] + d["err"]; n_skip += d.get("skip", 0)
        if d["run"] > 0 and d["fail"] == 0 and d["err"] == 0:
            green.append(d["module"])
        else:
            bad.append({"module": d["module"], "run": d["run"], "failed": d["fail"] + d["err"], "why": why.get(d["module"], ""),
                        "bad": d.get("bad", [])[:4]})
    out = {"what": old.get("what", "every test module of the reference run with B200Backend (host logic, CPU) in PandaPowerBackend's place"),
           "modules": len(mods), "fully_green": len(green), "with_failures": len(bad), "not_importable_here": len(noimp),
           "tests_run": n_run, "tests_failed_or_errored": n_bad, "tests_skipped_by_the_reference": n_skip,
           "green_modules": green, "modules_with_failures": bad, "not_importable": noimp}
    json.dump(out, open(OUT, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("modules", "fully_green", "with_failures", "not_importable_here", "tests_run", "tests_failed_or_errored")}))
    newly_bad = sorted(set(b["module"] for b in bad) - set(why))
    if newly_bad:
        print("NOT green any more (or new):", newly_bad)


if __name__ == "__main__":
    main()
