#!/bin/bash
# round 2, GPU call 8 (1 GPU): validation of the static-blob staging before the 8-GPU run: touched tests, variants, bench
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== pytest"; timeout 1800 python -m pytest tests/test_block_kernel_gpu.py tests/test_redo_gpu.py tests/test_engine_random_gpu.py tests/test_protections_gpu.py tests/test_n1_gpu.py -q -m gpu --tb=short > gpurun_out/pytest_failing.txt 2>&1; tail -6 gpurun_out/pytest_failing.txt
run() { local name=$1 batch=$2; shift 2
  env "$@" timeout 300 python bench.py --batch $batch --steps 60 --warmup 5 --no-cpu --e2e-groups 0 2>gpurun_out/var_err.txt | tail -1 > gpurun_out/var.json
  python -c "import json;d=json.load(open('gpurun_out/var.json'));print('$name B=$batch',round(d['value']/1e6,2),'M/s',round(1e3*d['ms_per_step'],2),'us',d['config']['launch'],d['parity_check']['ok'])" 2>/dev/null || tail -4 gpurun_out/var_err.txt
}
for B in 4096 16384 65536; do
  run "default(stage plan+static,wpc4,lockstep,pdl)" $B A=1
  run "wpc2" $B B200PF_BLOCK_WPC=2
  run "noredo" $B B200PF_NO_REDO=1
done
echo "== ncu default block kernel, batch 4096"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_block -s 6 -c 1 -o gpurun_out/prof_a python bench.py --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu1.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_a.ncu-rep gpurun_out/round2_ncu_block_case14 "pf_kernel_block<8,1,WPC=4,staged plan + static arrays (TMA),lockstep> l2rpn_case14_sandbox batch 4096, bench.py step (ncu --set full --clock-control none)" 4096 "planned_block:case14:T8"
rm -f gpurun_out/prof_a.ncu-rep
echo "== bench default"; timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1.json
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['lockstep_value'],d['parity_check']['ok'],d['cpu_baseline']['value'],d['roofline']['frac'])"
du -sh gpurun_out
