#!/bin/bash
# seuclidean / register row sort / warps-per-CTA variants: parity first, then A/B timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_metrics_gpu.py tests/test_select_gpu.py tests/test_smc_gpu.py tests/test_distance_gpu.py tests/test_store_gpu.py -x -q > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -5 gpurun_out/r2p_pytest.log
ELFI_B200_FUSED_WARPS=16 timeout 300 python -m pytest tests/test_smc_gpu.py -x -q -k "fused" > gpurun_out/r2p_pytest_w16.log 2>&1; tail -2 gpurun_out/r2p_pytest_w16.log
ELFI_B200_NESTED_WARPS=12 timeout 300 python -m pytest tests/test_distance_gpu.py -x -q > gpurun_out/r2p_pytest_n12.log 2>&1; tail -2 gpurun_out/r2p_pytest_n12.log
ELFI_B200_NESTED_WARPS=16 timeout 300 python -m pytest tests/test_distance_gpu.py -x -q > gpurun_out/r2p_pytest_n16.log 2>&1; tail -2 gpurun_out/r2p_pytest_n16.log
{
TAG=default timeout 300 python scripts/bench_r2_variants.py
TAG=rowsort_smem WHAT=rowsort ELFI_B200_ROWSORT_SMEM=1 timeout 300 python scripts/bench_r2_variants.py
TAG=warps8 WHAT=fused,nested ELFI_B200_FUSED_WARPS=8 ELFI_B200_NESTED_WARPS=8 timeout 300 python scripts/bench_r2_variants.py
TAG=warps12 WHAT=fused,nested ELFI_B200_FUSED_WARPS=12 ELFI_B200_NESTED_WARPS=12 timeout 300 python scripts/bench_r2_variants.py
TAG=warps16 WHAT=fused,nested ELFI_B200_FUSED_WARPS=16 ELFI_B200_NESTED_WARPS=16 timeout 300 python scripts/bench_r2_variants.py
} > gpurun_out/r2p_variants.jsonl 2> gpurun_out/r2p_variants.err
cat gpurun_out/r2p_variants.jsonl
tail -3 gpurun_out/r2p_variants.err
