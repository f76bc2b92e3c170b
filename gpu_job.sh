#!/bin/bash
mkdir -p gpurun_out
echo "== protections" ; timeout 900 python -m pytest tests/test_protections_gpu.py -x -q -m gpu 2>&1 | tail -15
echo "== pytest gpu (rest)" ; timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_protections_gpu.py 2>&1 | tail -5
echo "== bench"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu 2>gpurun_out/bench_err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
tail -3 gpurun_out/bench_err.txt
