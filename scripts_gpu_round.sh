#!/bin/bash
# one GPU call: gpu tests, smoke, bench (ours + reference arm), secondary configs, ncu launch list + full captures of the
# planned kernel (case14 bench step, 118-substation grid); artefacts land in gpurun_out/ (copy what is judged to profiles/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "== bench ours"; timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_n1.json
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['lockstep_value'],d['config']['kernel'],d['roofline']['frac'],d['cpu_baseline']['value'])"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 40 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_ref.json; cut -c1-200 gpurun_out/bench_ref.json
echo "== other configs"; timeout 900 python scripts/bench_configs.py > gpurun_out/other_configs.json 2>gpurun_out/other_err.txt; tail -2 gpurun_out/other_err.txt
if [ "$1" != "nocu" ]; then
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu planned case14"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_sparse -s 6 -c 1 -o gpurun_out/prof_planned_case14 python bench.py --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu1.log 2>&1
echo "== ncu planned 118"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_sparse -s 1 -c 1 -o gpurun_out/prof_planned_118 python scripts/profile_generic.py l2rpn_wcci_2022_dev 2048 > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log
fi
