#!/bin/bash
# round 2, 8-GPU call: BASELINE configs[1] (case14, 4096 per GPU, weak) and configs[4] (118 substations, 8192 in total, strong)
# at N = 1, 2, 4, 8; result collection: device ring + one asynchronous NCCL gather per 64 steps (default); one N = 8 run of the kernel-stored (CUDA IPC) alternative
mkdir -p gpurun_out
tr() { local n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $n "$@" 2>>gpurun_out/scale_err.txt | tail -1; }
show() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('$1', 'N=%d'%d['n_gpus'], round(d['value']/1e6,2),'M/s', round(1e3*d['ms_per_step'],2),'us/step', 'e2e',round(d['e2e']['value']/1e6,2), d['config']['result_collection'][:40], d['config']['collected_equals_results'], d['spread']['median'])"; }
nvidia-smi topo -m > gpurun_out/scale8_topo.txt 2>&1
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu 2>>gpurun_out/scale_err.txt | tail -1 | tee gpurun_out/scale8_case14_n1.json | show case14
for n in 2 4 8; do tr $n --steps 200 --warmup 10 --no-cpu | tee gpurun_out/scale8_case14_n$n.json | show case14; done
timeout 600 python bench.py --workload wcci --steps 60 --warmup 5 --no-cpu 2>>gpurun_out/scale_err.txt | tail -1 | tee gpurun_out/scale8_wcci_n1.json | show wcci
for n in 2 4 8; do tr $n --workload wcci --steps 60 --warmup 5 --no-cpu | tee gpurun_out/scale8_wcci_n$n.json | show wcci; done
tr 8 --steps 200 --warmup 10 --no-cpu --collect p2p --e2e-groups 0 | tee gpurun_out/scale8_case14_n8_p2p.json | show case14-p2p
tr 8 --batch 8192 --steps 100 --warmup 10 --no-cpu --e2e-groups 0 | tee gpurun_out/scale8_case14_b8192_n8.json | show case14-b8192
grep -v "^$" gpurun_out/scale_err.txt | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -8 | cut -c1-200
