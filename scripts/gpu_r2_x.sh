#!/bin/bash
# after the one-call top-n merge: full GPU suite on a 2-GPU box (includes the NCCL tests), then
# the 2-rank parity check and bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2x_pytest_gpu.log; cat gpurun_out/r2x_pytest_gpu.log
bash scripts/gpu_r2_b.sh 2
