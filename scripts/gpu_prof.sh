#!/bin/bash
# ncu --set full captures of the hot kernels (one or two launches each)
mkdir -p gpurun_out
for t in ${TARGETS:-nested summ gm gp}; do
  case $t in
    nested) k='regex:rowstream_kernel';;
    summ)   k='regex:rowstream_kernel';;
    gm)     k='regex:gm_pdf_kernel';;
    gp)     k='regex:gemm_nt_dmma';;
  esac
  skip=1; cnt=1
  if [ $t == gp ]; then skip=94; cnt=2; fi
  if [ $t == summ ]; then skip=1; cnt=4; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k "$k" -s $skip -c $cnt -f \
      -o gpurun_out/prof_$t python scripts/prof_targets.py $t > gpurun_out/ncu_$t.log 2>&1
  tail -2 gpurun_out/ncu_$t.log
done
ls -la gpurun_out/*.ncu-rep
