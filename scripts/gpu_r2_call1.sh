#!/bin/bash
# round 2, GPU call 1: full GPU test-suite (safety net, 20k-state stress, reference's extended suites through CUDA), bench
# with / without the safety net, config-5 workload at N=1, CPU arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt; lscpu | head -20 >> gpurun_out/gpu.txt
echo "== pytest gpu"; timeout 2400 python -m pytest tests -x -q -m gpu --durations=15 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "== bench ours"; timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_err.txt
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['value'],d['ms_per_step'],d['spread'],d['e2e']['value'],d['e2e']['lockstep_value'],d['parity_check'],d['gpu_launches'],d['cpu_baseline']['value'],d['cpu_baseline']['sample'])"
echo "== bench ours, safety net off"; timeout 600 python bench.py --steps 200 --warmup 10 --no-redo --no-cpu --e2e-groups 0 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1_noredo.json
python -c "import json;d=json.load(open('gpurun_out/bench_n1_noredo.json'));print(d['value'],d['ms_per_step'],d['spread'])"
echo "== bench wcci N=1"; timeout 600 python bench.py --workload wcci --steps 50 --warmup 5 --no-cpu 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_wcci_n1.json
python -c "import json;d=json.load(open('gpurun_out/bench_wcci_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_check'],d['config']['launch'])"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref.json; python -c "import json;d=json.load(open('gpurun_out/bench_ref.json'));print(d['value'],d['spread'],d['cpu_baseline']['sample'])"
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref2.json; python -c "import json;d=json.load(open('gpurun_out/bench_ref2.json'));print(d['value'],d['spread'])"
