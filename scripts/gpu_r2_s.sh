#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_gp_fit.py 2>&1 | grep "max err" > gpurun_out/r2s_gp.log
timeout 900 python -m pytest tests/test_gp_gpu.py tests/test_bolfi_gpu.py -x -q 2>&1 | tail -3 >> gpurun_out/r2s_gp.log
for sb in 0 100 200; do
  TAG="small_below=$sb" ELFI_B200_GEMM_SMALL_BELOW=$sb timeout 120 python scripts/time_gp_fit.py 2>&1 | tail -1 >> gpurun_out/r2s_gp.log
done
for sb in 0 100; do
  echo "bench_kernels small_below=$sb" >> gpurun_out/r2s_gp.log
  ELFI_B200_GEMM_SMALL_BELOW=$sb timeout 600 python scripts/bench_kernels.py 2>/dev/null | grep -i "GP \|gp_\|LCBSC" | cut -c1-160 >> gpurun_out/r2s_gp.log
done
cat gpurun_out/r2s_gp.log
