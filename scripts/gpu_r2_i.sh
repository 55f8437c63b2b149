#!/bin/bash
# round 2, call I (1 GPU): mixed-precision mixture density: speed and accuracy against the fp64 path
mkdir -p gpurun_out
TAG=gm_fp64 timeout 600 python scripts/bench_r2_kernels.py > gpurun_out/r2i_fp64.log 2>&1; grep gm_logpdf gpurun_out/r2i_fp64.log | cut -c1-300
TAG=gm_mixed ELFI_B200_GM_MODE=mixed timeout 600 python scripts/bench_r2_kernels.py > gpurun_out/r2i_mixed.log 2>&1; grep gm_logpdf gpurun_out/r2i_mixed.log | cut -c1-300
python scripts/gm_mode_compare.py gm_fp64 gm_mixed | tee gpurun_out/r2i_gm_accuracy.json
ELFI_B200_GM_MODE=mixed timeout 600 python -m pytest tests/test_smc_gpu.py tests/test_samplers_gpu.py tests/test_throughput_gpu.py -m gpu -q 2>&1 | tail -12
