#!/bin/bash
# round 2, 2-GPU call: scaling of the DoNothing bench with kernel-stored result collection over NVLink (CUDA IPC) vs the NCCL ring
# gather; config 5 (118 substations, batch 8192 in total) at N = 1, 2
mkdir -p gpurun_out
tr() { local n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $n "$@" 2>>gpurun_out/scale_err.txt | tail -1; }
show() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('$1', 'N=%d'%d['n_gpus'], round(d['value']/1e6,2),'M/s', round(1e3*d['ms_per_step'],2),'us/step', 'e2e',round(d['e2e']['value']/1e6,2), d['config']['result_collection'][:60], d['config']['collected_equals_results'])"; }
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu 2>>gpurun_out/scale_err.txt | tail -1 | tee gpurun_out/scale_case14_n1.json | show case14
tr 2 --steps 200 --warmup 10 --no-cpu | tee gpurun_out/scale_case14_n2_p2p.json | show case14-p2p
tr 2 --steps 200 --warmup 10 --no-cpu --collect nccl | tee gpurun_out/scale_case14_n2_nccl.json | show case14-nccl
timeout 600 python bench.py --workload wcci --steps 60 --warmup 5 --no-cpu 2>>gpurun_out/scale_err.txt | tail -1 | tee gpurun_out/scale_wcci_n1.json | show wcci
tr 2 --workload wcci --steps 60 --warmup 5 --no-cpu | tee gpurun_out/scale_wcci_n2.json | show wcci
tail -5 gpurun_out/scale_err.txt
