#!/bin/bash
mkdir -p gpurun_out
{
for pv in unrolled compact; do
  for sk in 0 1 2 3; do
    TAG="panel=$pv skip=$sk" ELFI_B200_GP_PANEL=$pv ELFI_B200_GP_SKIP=$sk timeout 120 python scripts/time_gp_fit.py 2>&1 | tail -1
  done
done
} > gpurun_out/r2r_gp_variants.log
cat gpurun_out/r2r_gp_variants.log
