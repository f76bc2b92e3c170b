#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== bench"; timeout 300 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.txt | tail -1 | tee gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['launch'], 'e2e', d['e2e']['value'], d['cpu_baseline'], d['clocks'])"
tail -3 gpurun_out/bench_err.txt
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 40 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-400
