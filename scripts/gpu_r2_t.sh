#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gp_gpu.py tests/test_bolfi_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/r2t_gp.log
for few in 0 1; do
  echo "bench_kernels GP_FEW=$few" >> gpurun_out/r2t_gp.log
  ELFI_B200_GP_FEW=$few timeout 600 python scripts/bench_kernels.py 2>/dev/null | grep -i "GP \|gp_\|LCBSC" | cut -c1-170 >> gpurun_out/r2t_gp.log
done
cat gpurun_out/r2t_gp.log
