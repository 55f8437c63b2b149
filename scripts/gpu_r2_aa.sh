#!/bin/bash
# round 2, call AA (1 GPU): full GPU suite after the model/store/bo rewrite + mean/var row-group A/B + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2aa_pytest_gpu.log; cat gpurun_out/r2aa_pytest_gpu.log
timeout 300 python scripts/time_meanvar.py > gpurun_out/r2aa_meanvar.jsonl 2> gpurun_out/r2aa_meanvar.err; cat gpurun_out/r2aa_meanvar.jsonl; tail -3 gpurun_out/r2aa_meanvar.err
timeout 240 ncu --set full --clock-control none --import-source on -k regex:meanvar_rowgroup -c 1 -f -o gpurun_out/r2aa_meanvar_rg python scripts/time_meanvar.py --once > gpurun_out/r2aa_ncu.log 2>&1; tail -2 gpurun_out/r2aa_ncu.log
