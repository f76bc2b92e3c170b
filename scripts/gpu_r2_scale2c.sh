#!/bin/bash
# 2-GPU check of bench.py after the RingCollector refactoring (default collection + p2p option)
mkdir -p gpurun_out
tr() { local n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $n "$@" 2>>gpurun_out/scale_err.txt | tail -1; }
show() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('$1', 'N=%d'%d['n_gpus'], round(d['value']/1e6,2),'M/s', round(1e3*d['ms_per_step'],2),'us/step', d['config']['result_collection'][:40], d['config']['collected_equals_results'], d['parity_check']['ok'])"; }
tr 2 --steps 150 --warmup 10 --no-cpu --e2e-groups 0 | tee gpurun_out/scale2c_case14_n2.json | show case14
tr 2 --steps 150 --warmup 10 --no-cpu --e2e-groups 0 --collect p2p | tee gpurun_out/scale2c_case14_n2_p2p.json | show case14-p2p
grep -v "^$" gpurun_out/scale_err.txt | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -6 | cut -c1-200
