#!/bin/bash
# round 2, GPU call 3: failing tests with tracebacks, TMA-staging variants, ncu captures summarised ON THE BOX (the reports are
# too big to travel back: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== pytest (files that failed)"; timeout 1800 python -m pytest tests/test_block_kernel_gpu.py tests/test_redo_gpu.py tests/test_batched_env.py tests/test_protections_gpu.py -q -m gpu --tb=short > gpurun_out/pytest_failing.txt 2>&1; tail -12 gpurun_out/pytest_failing.txt
echo "== staged variants"
for v in "8 1 4 1" "4 2 4 1"; do set -- $v
  B200PF_BLOCK_T=$1 B200PF_BLOCK_U=$2 B200PF_BLOCK_WPC=$3 B200PF_BLOCK_STAGE=$4 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --e2e-groups 0 2>gpurun_out/stage_err_$1.txt | tail -1 > gpurun_out/bench_T$1_U$2_W$3_S$4.json
  tail -5 gpurun_out/stage_err_$1.txt
  python -c "import json,sys;d=json.load(open('gpurun_out/bench_T$1_U$2_W$3_S$4.json'));print('T$1 U$2 WPC$3 STAGE$4',round(d['value']/1e6,2),'M/s',round(1e3*d['ms_per_step'],2),'us',d['config']['launch'])"
done
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_block.csv python bench.py --steps 20 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/bench_under_ncu.log 2>&1; grep -c pf_kernel gpurun_out/launches_block.csv
echo "== ncu block case14 (default T=8 U=1)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_block -s 6 -c 1 -o gpurun_out/prof_a python bench.py --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu1.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_a.ncu-rep gpurun_out/round2_ncu_block_case14 "pf_kernel_block<8,1> l2rpn_case14_sandbox batch 4096, bench.py step (ncu --set full --clock-control none)" 4096 "planned_block:case14:T8"
ncu -i gpurun_out/prof_a.ncu-rep --page source --csv --print-source cuda > gpurun_out/round2_ncu_block_case14_source.csv 2>/dev/null; rm -f gpurun_out/prof_a.ncu-rep
echo "== ncu block case14 batch 65536"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_block -s 6 -c 1 -o gpurun_out/prof_b python bench.py --batch 65536 --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu2.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_b.ncu-rep gpurun_out/round2_ncu_block_case14_b65536 "pf_kernel_block<8,1> l2rpn_case14_sandbox batch 65536 (ncu --set full --clock-control none)" 65536 "planned_block:case14:T8:b65536"
rm -f gpurun_out/prof_b.ncu-rep
echo "== ncu redo kernel"
timeout 600 ncu --set full --clock-control none -k regex:pf_kernel_redo -s 6 -c 1 -o gpurun_out/prof_c python bench.py --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu3.log 2>&1
ncu -i gpurun_out/prof_c.ncu-rep --page details 2>/dev/null | head -60 > gpurun_out/round2_ncu_redo_details.txt; rm -f gpurun_out/prof_c.ncu-rep
du -sh gpurun_out
