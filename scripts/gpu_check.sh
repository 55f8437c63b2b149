#!/bin/bash
# Run on the GPU box through gpurun: parity tests, bench, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" == "prof" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowstream -s 4 -c 2 \
      -f -o gpurun_out/prof_dist python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
