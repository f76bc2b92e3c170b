#!/bin/bash
# round 2, closing call: full GPU suite on the final tree (bus count as host code of the library), config 3 with the bus cap following the topology
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== pytest gpu all"; timeout 600 python -m pytest tests -q -m gpu --tb=short --durations=3 > gpurun_out/pytest_gpu_final4.txt 2>&1; tail -6 gpurun_out/pytest_gpu_final4.txt
echo "== config 3"; timeout 300 python scripts/bench_config3.py --steps 30 > gpurun_out/config3.json 2> gpurun_out/config3.log; python - <<'PY'
import json
d = json.load(open("gpurun_out/config3.json"))
for k in ("policy2", "policy1", "policy0", "policy1_every_slot"):
    r = d[k]; print(k, round(r["env_step_per_s_steady"]), round(r["ms_per_step_steady_median"], 3), round(r["ms_per_step_min"], 3), r["launch"], r["bus_cap"])
PY
