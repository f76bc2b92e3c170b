#!/bin/bash
# 2-GPU check of the completion-flag collection (no collective in the stepping loop)
mkdir -p gpurun_out
tr() { local n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $n "$@" 2>>gpurun_out/scale_err.txt | tail -1; }
show() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('$1', 'N=%d'%d['n_gpus'], round(d['value']/1e6,2),'M/s', round(1e3*d['ms_per_step'],2),'us/step', 'e2e',round(d['e2e']['value']/1e6,2), d['config']['result_collection'][:40], d['config']['collected_equals_results'], d['spread']['median'], d['spread']['tail_ms'])"; }
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --e2e-groups 0 2>>gpurun_out/scale_err.txt | tail -1 | tee gpurun_out/scale2b_case14_n1.json | show case14
tr 2 --steps 200 --warmup 10 --no-cpu --e2e-groups 0 | tee gpurun_out/scale2b_case14_n2.json | show case14-p2p-flag
tr 2 --steps 200 --warmup 10 --no-cpu --e2e-groups 0 --collect p2p | tee gpurun_out/scale2b_case14_n2_p2p.json | show case14-p2p
timeout 300 python -m pytest tests/test_protections_gpu.py tests/test_redo_gpu.py -q -m gpu 2>&1 | tail -2
grep -v "^$" gpurun_out/scale_err.txt | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -6 | cut -c1-200
