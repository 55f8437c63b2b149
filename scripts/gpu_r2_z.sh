#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_throughput_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/r2z_pytest.log; cat gpurun_out/r2z_pytest.log
ELFI_B200_SIM_GAUSS_TREE=1 timeout 600 python -m pytest tests/test_throughput_gpu.py -x -q -k gauss 2>&1 | tail -2
{
TAG=default timeout 300 python scripts/time_sims.py
TAG=two_pass_tree ELFI_B200_SIM_GAUSS_TREE=1 ELFI_B200_SIM_MA2_TREE=1 timeout 300 python scripts/time_sims.py
} 2>&1 | tee gpurun_out/r2z_time_sims.log
