#!/bin/bash
# round 2, GPU call 4: the tests that failed, register-budget / staging variants of the block kernel at three batch sizes
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== pytest (files that failed)"; timeout 1800 python -m pytest tests/test_block_kernel_gpu.py tests/test_redo_gpu.py tests/test_batched_env.py tests/test_protections_gpu.py -q -m gpu --tb=short > gpurun_out/pytest_failing.txt 2>&1; tail -8 gpurun_out/pytest_failing.txt
echo "== variants"
run() { # name, batch, env...
  local name=$1 batch=$2; shift 2
  env "$@" timeout 300 python bench.py --batch $batch --steps 60 --warmup 5 --no-cpu --e2e-groups 0 2>gpurun_out/var_err.txt | tail -1 > gpurun_out/var.json
  python -c "import json;d=json.load(open('gpurun_out/var.json'));print('$name B=$batch',round(d['value']/1e6,2),'M/s',round(1e3*d['ms_per_step'],2),'us',d['config']['launch'],d['parity_check']['ok'])" 2>/dev/null || tail -3 gpurun_out/var_err.txt
}
for B in 4096 16384 65536; do
  run "free minb8" $B B200PF_BLOCK_MINB=8 B200PF_BLOCK_UNI=0
  run "free minb20" $B B200PF_BLOCK_MINB=20 B200PF_BLOCK_UNI=0
  run "lockstep minb8" $B B200PF_BLOCK_MINB=8
  run "lockstep minb16" $B B200PF_BLOCK_MINB=16
  run "lockstep minb20" $B B200PF_BLOCK_MINB=20
  run "wpc4" $B B200PF_BLOCK_WPC=4
  run "wpc4+stage" $B B200PF_BLOCK_STAGE=1
  run "noredo" $B B200PF_NO_REDO=1
done
run "scalar" 4096 B200PF_BLOCK=0
run "scalar" 65536 B200PF_BLOCK=0
echo "== ncu block case14 minb20 batch 65536"
B200PF_BLOCK_MINB=20 B200PF_BLOCK_UNI=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_block -s 6 -c 1 -o gpurun_out/prof_b python bench.py --batch 65536 --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu2.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_b.ncu-rep gpurun_out/round2_ncu_block_case14_minb20_b65536 "pf_kernel_block<8,1,MINB=20> l2rpn_case14_sandbox batch 65536 (ncu --set full --clock-control none)" 65536 "planned_block:case14:T8:minb20:b65536"
rm -f gpurun_out/prof_b.ncu-rep
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_block.csv python bench.py --steps 20 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/bench_under_ncu.log 2>&1; grep -c pf_kernel gpurun_out/launches_block.csv
du -sh gpurun_out
