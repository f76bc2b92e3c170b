#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/bench_configs.py > gpurun_out/other_configs.json 2> gpurun_out/other_err.txt; cat gpurun_out/other_configs.json; tail -3 gpurun_out/other_err.txt
