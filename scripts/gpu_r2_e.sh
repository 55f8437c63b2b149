#!/bin/bash
# round 2, call E (1 GPU): GP tests after the fit rework, kernel timings (+ variants), configs 3 / 5 at 1 GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gp_gpu.py tests/test_bolfi_gpu.py tests/test_summaries_gpu.py tests/test_smc_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2e_pytest.log; tail -6 gpurun_out/r2e_pytest.log
TAG=default timeout 600 python scripts/bench_r2_kernels.py > gpurun_out/r2e_kernels_default.log 2>&1; grep -E '"name"' gpurun_out/r2e_kernels_default.log | cut -c1-260
TAG=termwise ELFI_B200_SUMM_TERMWISE=1 ELFI_B200_MEANVAR_TWO_SWEEPS=1 timeout 600 python scripts/bench_r2_kernels.py > gpurun_out/r2e_kernels_termwise.log 2>&1; grep -E 'meanvar|autocov' gpurun_out/r2e_kernels_termwise.log | cut -c1-260
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2e_bench_kernels.log 2>&1; grep -E "gp_|rank|dist_euclid" gpurun_out/r2e_bench_kernels.log | cut -c1-260
bash scripts/gpu_r2_configs.sh 1
