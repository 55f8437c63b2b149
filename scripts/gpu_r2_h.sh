#!/bin/bash
# round 2, call H (2 GPUs): merge test, bench at 1 and 2 ranks after the one-call step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_merge_gpu.py tests/test_gp_gpu.py -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench_g1.json 2> gpurun_out/r2h_bench_g1.err; echo "g1 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench_g1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','host_launch_ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['smc_ma2']['seconds'])
PY
bash scripts/gpu_r2_b.sh 2
