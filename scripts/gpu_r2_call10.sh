#!/bin/bash
# round 2, GPU call 10 (1 GPU): full suite, default bench, launch list + ncu of the final default kernel, other workloads
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== pytest gpu all"; timeout 2400 python -m pytest tests -q -m gpu --tb=short --durations=5 > gpurun_out/pytest_gpu_full.txt 2>&1; tail -8 gpurun_out/pytest_gpu_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench default"; timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1.json
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['lockstep_value'],d['parity_check']['ok'],d['cpu_baseline']['value'],d['roofline']['frac'],d['config']['launch'])"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref.json; python -c "import json;d=json.load(open('gpurun_out/bench_ref.json'));print(d['value'],d['spread'])"
run() { local name=$1 batch=$2; shift 2
  env "$@" timeout 300 python bench.py --batch $batch --steps 60 --warmup 5 --no-cpu --e2e-groups 0 2>gpurun_out/var_err.txt | tail -1 > gpurun_out/var.json
  python -c "import json;d=json.load(open('gpurun_out/var.json'));print('$name B=$batch',round(d['value']/1e6,2),'M/s',round(1e3*d['ms_per_step'],2),'us',d['config']['launch'],d['parity_check']['ok'])" 2>/dev/null || tail -4 gpurun_out/var_err.txt
}
for B in 2048 4096 8192 16384 32768 65536; do run "default" $B A=1; done
run "scalar planned kernel" 4096 B200PF_BLOCK=0
run "scalar planned kernel" 65536 B200PF_BLOCK=0
echo "== bench wcci"; timeout 600 python bench.py --workload wcci --steps 60 --warmup 5 --no-cpu 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_wcci_n1.json
python -c "import json;d=json.load(open('gpurun_out/bench_wcci_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_check']['ok'],d['roofline']['frac'],d['config']['launch'])"
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_block.csv python bench.py --steps 20 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/bench_under_ncu.log 2>&1; grep -c pf_kernel gpurun_out/launches_block.csv
echo "== ncu default block kernel, batch 4096"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_block -s 6 -c 1 -o gpurun_out/prof_a python bench.py --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu1.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_a.ncu-rep gpurun_out/round2_ncu_block_case14 "pf_kernel_block<8,1,WPC=2,staged plan + static arrays (TMA),lockstep> l2rpn_case14_sandbox batch 4096, bench.py step (ncu --set full --clock-control none)" 4096 "planned_block:case14:T8"
rm -f gpurun_out/prof_a.ncu-rep
echo "== ncu default block kernel, batch 65536"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_block -s 6 -c 1 -o gpurun_out/prof_b python bench.py --batch 65536 --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu2.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_b.ncu-rep gpurun_out/round2_ncu_block_case14_b65536 "pf_kernel_block<8,1,WPC=4,staged plan + static arrays (TMA),lockstep> l2rpn_case14_sandbox batch 65536 (ncu --set full --clock-control none)" 65536 "planned_block:case14:T8:b65536"
rm -f gpurun_out/prof_b.ncu-rep
echo "== ncu scalar kernel 118 substations, batch 2048"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_sparse -s 3 -c 1 -o gpurun_out/prof_c python bench.py --workload wcci --batch 2048 --steps 6 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu3.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_c.ncu-rep gpurun_out/round2_ncu_sparse_118 "pf_kernel_sparse<64> l2rpn_wcci_2022_dev batch 2048 (ncu --set full --clock-control none)" 2048 "planned_sparse:wcci:T64"
rm -f gpurun_out/prof_c.ncu-rep
du -sh gpurun_out
