#!/bin/bash
# round 2, call AB (1 GPU): GP tests after the interleaved sub-tile / triangular-skip GEMM change, predict A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gp_gpu.py tests/test_bolfi_gpu.py tests/test_summaries_gpu.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2ab_pytest_gp.log; cat gpurun_out/r2ab_pytest_gp.log
timeout 400 python scripts/time_gp_predict.py > gpurun_out/r2ab_gp_predict.jsonl 2> gpurun_out/r2ab_gp_predict.err; cat gpurun_out/r2ab_gp_predict.jsonl; tail -3 gpurun_out/r2ab_gp_predict.err
timeout 200 python scripts/time_meanvar.py --child 2>/dev/null | grep -E "x62|x50" > gpurun_out/r2ab_meanvar62.jsonl; cat gpurun_out/r2ab_meanvar62.jsonl
