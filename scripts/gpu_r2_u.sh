#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gp_gpu.py tests/test_bolfi_gpu.py tests/test_store_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/r2u_gp.log
timeout 600 python scripts/bench_kernels.py 2>/dev/null | grep -i "GP \|gp_\|LCBSC" | cut -c1-170 >> gpurun_out/r2u_gp.log
cp gpurun_out/kernels.json gpurun_out/r2u_kernels.json
cat gpurun_out/r2u_gp.log
