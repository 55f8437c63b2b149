#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2v_gp.log
for bk in 1 0; do
  echo "BK32=$bk" >> gpurun_out/r2v_gp.log
  ELFI_B200_GEMM_BK32=$bk timeout 900 python -m pytest tests/test_gp_gpu.py -x -q 2>&1 | tail -2 >> gpurun_out/r2v_gp.log
  ELFI_B200_GEMM_BK32=$bk timeout 600 python scripts/bench_kernels.py 2>/dev/null | grep -i "K12\|K10\|m=128" | cut -c1-200 >> gpurun_out/r2v_gp.log
done
cat gpurun_out/r2v_gp.log
