#!/bin/bash
# round 2, last 1-GPU call: full GPU suite on the final tree (Simulator / Runner / bus-cap tests added), config 3 with the bus cap
# following the topology, safety-net cost on the AC/rows instantiation
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== pytest gpu all"; timeout 900 python -m pytest tests -q -m gpu --tb=short --durations=5 > gpurun_out/pytest_gpu_final3.txt 2>&1; tail -8 gpurun_out/pytest_gpu_final3.txt
echo "== config 3"; timeout 600 python scripts/bench_config3.py --steps 30 > gpurun_out/config3.json 2> gpurun_out/config3.log; tail -4 gpurun_out/config3.log | cut -c1-900
echo "== no redo"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu --no-redo --e2e-groups 0 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1_noredo.json
python -c "import json;d=json.load(open('gpurun_out/bench_n1_noredo.json'));print(d['value'],d['ms_per_step'])"
echo "== with redo"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu --e2e-groups 0 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1_redo.json
python -c "import json;d=json.load(open('gpurun_out/bench_n1_redo.json'));print(d['value'],d['ms_per_step'])"
