#!/bin/bash
# gpurun helper: run the GPU parity tests (optionally a subset) and keep the log.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q "$@" 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
