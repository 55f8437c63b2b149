#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_smc_gpu.py tests/test_reference_plugin.py tests/test_bolfi_gpu.py tests/test_model_gpu.py tests/test_samplers_gpu.py tests/test_merge_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2c_pytest.log; tail -30 gpurun_out/r2c_pytest.log
timeout 300 python scripts/prof_smc_host.py 1000000 125000 3 > gpurun_out/r2_prof_smc_host.txt 2>&1; head -120 gpurun_out/r2_prof_smc_host.txt | cut -c1-180
