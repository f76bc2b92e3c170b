#!/bin/bash
# round 2, GPU call 1: full GPU test-suite (block-planned kernel as default, safety net, 20k-state stress, reference's extended
# suites through CUDA), block-kernel sweep, bench with / without the safety net, config-5 workload at N=1, CPU arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt; lscpu | head -20 >> gpurun_out/gpu.txt
echo "== pytest gpu (block kernel tests first)"; timeout 900 python -m pytest tests/test_block_kernel_gpu.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_block.txt
echo "== sweep block case14"; timeout 600 python scripts/sweep_block.py case14 > gpurun_out/sweep_block_case14.json 2> gpurun_out/sweep_block_case14.log; tail -32 gpurun_out/sweep_block_case14.log
echo "== pytest gpu all"; timeout 2400 python -m pytest tests -q -m gpu --durations=12 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "== bench ours"; timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_err.txt
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['value'],d['ms_per_step'],d['spread'],d['e2e']['value'],d['e2e']['lockstep_value'],d['parity_check'],d['gpu_launches'],d['config']['launch'],d['cpu_baseline']['value'],d['cpu_baseline']['sample'])"
echo "== bench ours, safety net off"; timeout 600 python bench.py --steps 200 --warmup 10 --no-redo --no-cpu --e2e-groups 0 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1_noredo.json
python -c "import json;d=json.load(open('gpurun_out/bench_n1_noredo.json'));print(d['value'],d['ms_per_step'],d['spread'])"
echo "== bench ours, scalar planned kernel"; B200PF_BLOCK=0 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --e2e-groups 0 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1_scalar.json
python -c "import json;d=json.load(open('gpurun_out/bench_n1_scalar.json'));print(d['value'],d['ms_per_step'],d['spread'])"
echo "== sweep block n36 wcci"; timeout 900 python scripts/sweep_block.py n36 wcci > gpurun_out/sweep_block_big.json 2> gpurun_out/sweep_block_big.log; tail -40 gpurun_out/sweep_block_big.log
echo "== bench wcci N=1"; timeout 600 python bench.py --workload wcci --steps 50 --warmup 5 --no-cpu 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_wcci_n1.json
python -c "import json;d=json.load(open('gpurun_out/bench_wcci_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_check'],d['config']['launch'])"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref.json; python -c "import json;d=json.load(open('gpurun_out/bench_ref.json'));print(d['value'],d['spread'],d['cpu_baseline']['sample'])"
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref2.json; python -c "import json;d=json.load(open('gpurun_out/bench_ref2.json'));print(d['value'],d['spread'])"
echo "== config 3"; timeout 900 python scripts/bench_config3.py --steps 30 > gpurun_out/config3.json 2> gpurun_out/config3.log; tail -3 gpurun_out/config3.log | cut -c1-600
echo "== single env latency"; NB_TS=300 timeout 900 python scripts/single_env_latency.py > gpurun_out/single_env_latency.json 2> gpurun_out/single_env_latency.log; grep l2rpn gpurun_out/single_env_latency.log | cut -c1-500
