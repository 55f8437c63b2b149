#!/bin/bash
# Final round check on one B200: smoke, full GPU tests, bench (both arms), ncu launch list of the
# timed step, fresh kernel timings.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rowstream|compact_mask" -s 10 -c 40 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python scripts/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; tail -3 gpurun_out/bench_kernels.log | cut -c1-200
