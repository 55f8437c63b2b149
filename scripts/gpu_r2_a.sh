#!/bin/bash
# round 2, call A (1 GPU): GPU parity suite, bench reference arm, bench our arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2a_pytest_gpu.log; tail -8 gpurun_out/r2a_pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"; cut -c1-3000 gpurun_out/r2a_bench_ref.json; tail -5 gpurun_out/r2a_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "ours rc=$?"; cut -c1-6000 gpurun_out/r2a_bench.json; tail -15 gpurun_out/r2a_bench.err
