#!/bin/bash
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== bench"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu 2>gpurun_out/bench_err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
tail -3 gpurun_out/bench_err.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --batch 148 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('b148', d['value'], d['ms_per_step'])"
