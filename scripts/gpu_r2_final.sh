#!/bin/bash
# round 2, final 1-GPU call: full GPU suite, smoke, default bench (artefacts of the final tree)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== pytest gpu all"; timeout 1500 python -m pytest tests -q -m gpu --tb=short --durations=5 > gpurun_out/pytest_gpu_full.txt 2>&1; tail -8 gpurun_out/pytest_gpu_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench default"; timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1.json
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['lockstep_value'],d['parity_check']['ok'],d['cpu_baseline']['value'],d['roofline']['frac'],d['roofline']['traffic'],d['config']['launch'])"
tail -3 gpurun_out/bench_err.txt
