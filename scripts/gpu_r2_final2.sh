#!/bin/bash
# round 2, closing 1-GPU call on the final tree: default bench (+cpu_baseline), reference arm, launch list, ncu of the default kernel
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== bench default"; timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1.json
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['lockstep_value'],d['parity_check']['ok'],d['cpu_baseline']['value'],d['roofline']['frac'],d['roofline']['traffic'],d['config']['launch'],d['clocks'])"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ref.json; python -c "import json;d=json.load(open('gpurun_out/bench_ref.json'));print(d['value'],d.get('spread'),d['cpu_baseline'])"
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_block.csv python bench.py --steps 20 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/bench_under_ncu.log 2>&1; grep -c pf_kernel gpurun_out/launches_block.csv
echo "== ncu default block kernel, batch 4096"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_block -s 6 -c 1 -o gpurun_out/prof_a python bench.py --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu1.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_a.ncu-rep gpurun_out/round2_ncu_block_case14 "pf_kernel_block<8,1,WPC=2,staged plan + static arrays (TMA),lockstep,MODE=1 (AC, chronics rows)> l2rpn_case14_sandbox batch 4096, bench.py step (ncu --set full --clock-control none)" 4096 "planned_block:case14:T8"
rm -f gpurun_out/prof_a.ncu-rep
echo "== bench wcci"; timeout 600 python bench.py --workload wcci --steps 60 --warmup 5 --no-cpu 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_wcci_n1.json
python -c "import json;d=json.load(open('gpurun_out/bench_wcci_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_check']['ok'],d['roofline']['frac'],d['config']['launch'])"
du -sh gpurun_out; tail -3 gpurun_out/bench_err.txt
