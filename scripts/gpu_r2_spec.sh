#!/bin/bash
# round 2: AC/rows-specialised block kernel (MODE=1) vs the generic one, per-instance reset; full GPU suite on the new tree
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== pytest gpu all"; timeout 900 python -m pytest tests -q -m gpu --tb=short --durations=5 > gpurun_out/pytest_gpu_spec.txt 2>&1; tail -8 gpurun_out/pytest_gpu_spec.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for rep in 1 2 3; do
  for spec in 1 0; do
    B200PF_BLOCK_SPEC=$spec timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu 2>>gpurun_out/bench_spec_err.txt | tail -1 > gpurun_out/bench_spec${spec}_r${rep}.json
    python -c "import json;d=json.load(open('gpurun_out/bench_spec${spec}_r${rep}.json'));print('spec',${spec},'rep',${rep},d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_check']['ok'] if 'parity_check' in d else None)"
  done
done
for spec in 1 0; do
  B200PF_BLOCK_SPEC=$spec timeout 300 python bench.py --steps 100 --warmup 10 --batch 65536 --no-cpu 2>>gpurun_out/bench_spec_err.txt | tail -1 > gpurun_out/bench_spec${spec}_b65536.json
  python -c "import json;d=json.load(open('gpurun_out/bench_spec${spec}_b65536.json'));print('spec',${spec},'b65536',d['value'],d['ms_per_step'])"
done
tail -3 gpurun_out/bench_spec_err.txt
