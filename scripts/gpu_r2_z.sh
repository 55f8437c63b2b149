#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_throughput_gpu.py tests/test_samplers_gpu.py tests/test_smc_gpu.py tests/test_merge_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/r2z_pytest.log; cat gpurun_out/r2z_pytest.log
timeout 300 python scripts/prof_rejection_torch.py > gpurun_out/r2z_prof_rejection.txt 2>&1; cut -c1-200 gpurun_out/r2z_prof_rejection.txt | grep -v "^---" | head -40
