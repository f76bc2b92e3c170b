#!/bin/bash
# one GPU call: gpu tests, smoke, bench (ours + reference arm), ncu launch list + full capture of the top kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
echo "== bench ours" ; timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.txt | tail -1 | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 40 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
if [ "$1" == "ncu" ]; then
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf_kernel -s 5 -c 2 -o gpurun_out/prof python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
fi
