#!/bin/bash
# the round-end sequence of gpu_r2_j.sh plus the per-kernel tables
bash scripts/gpu_r2_j.sh
timeout 900 python scripts/bench_kernels.py > gpurun_out/r2w_bench_kernels.log 2>&1; cp gpurun_out/kernels.json gpurun_out/r2w_kernels.json
grep -c name gpurun_out/r2w_bench_kernels.log
timeout 600 python scripts/bench_r2_kernels.py > gpurun_out/r2w_bench_r2_kernels.log 2>&1; tail -3 gpurun_out/r2w_bench_r2_kernels.log | cut -c1-200
