#!/bin/bash
# round 2, GPU call 2: full GPU test-suite with tracebacks, ncu launch list + full captures of the block kernel, CTA-shape / TMA
# staging variants, config 3 after the fixes
mkdir -p gpurun_out
echo "== pytest gpu all"; timeout 2400 python -m pytest tests -q -m gpu --tb=short --durations=12 > gpurun_out/pytest_gpu_full.txt 2>&1; tail -25 gpurun_out/pytest_gpu_full.txt
echo "== bench variants (device-timed only)"
for v in "8 1 1 0" "8 1 4 0" "8 1 4 1" "4 2 1 0" "4 2 4 0" "4 2 4 1"; do set -- $v
  B200PF_BLOCK_T=$1 B200PF_BLOCK_U=$2 B200PF_BLOCK_WPC=$3 B200PF_BLOCK_STAGE=$4 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --e2e-groups 0 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_T$1_U$2_W$3_S$4.json
  python -c "import json,sys;d=json.load(open('gpurun_out/bench_T$1_U$2_W$3_S$4.json'));print('T$1 U$2 WPC$3 STAGE$4',round(d['value']/1e6,2),'M/s',round(1e3*d['ms_per_step'],2),'us',d['config']['launch'])"
done
for B in 16384 65536; do B200PF_BLOCK_WPC=4 B200PF_BLOCK_STAGE=1 timeout 300 python bench.py --batch $B --steps 50 --warmup 5 --no-cpu --e2e-groups 0 2>>gpurun_out/bench_err.txt | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('staged batch',$B,round(d['value']/1e6,2),'M/s')"; timeout 300 python bench.py --batch $B --steps 50 --warmup 5 --no-cpu --e2e-groups 0 2>>gpurun_out/bench_err.txt | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('default batch',$B,round(d['value']/1e6,2),'M/s')"; done
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_block.csv python bench.py --steps 20 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/bench_under_ncu.log 2>&1; grep -c pf_kernel gpurun_out/launches_block.csv
echo "== ncu block case14 (default T=8 U=1)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_block -s 6 -c 1 -o gpurun_out/prof_block_case14_T8 python bench.py --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log | cut -c1-200
echo "== ncu block case14 batch 65536"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel_block -s 6 -c 1 -o gpurun_out/prof_block_case14_T8_B65536 python bench.py --batch 65536 --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log | cut -c1-200
echo "== ncu redo kernel"
timeout 600 ncu --set full --clock-control none -k regex:pf_kernel_redo -s 6 -c 1 -o gpurun_out/prof_redo python bench.py --steps 8 --warmup 3 --no-cpu --e2e-groups 0 > gpurun_out/ncu3.log 2>&1
echo "== config 3"; timeout 900 python scripts/bench_config3.py --steps 30 > gpurun_out/config3.json 2> gpurun_out/config3.log; tail -4 gpurun_out/config3.log | cut -c1-700
ls -la gpurun_out/*.ncu-rep
