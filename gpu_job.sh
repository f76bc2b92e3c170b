#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== bench"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu 2>gpurun_out/bench_err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['launch'], 'e2e', d['e2e']['value'])"
tail -3 gpurun_out/bench_err.txt
for b in 148 8192; do echo "== batch $b"; timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --batch $b 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['launch'])"; done
