#!/bin/bash
# round 2, call F (1 GPU): re-time the fused kernel / GEMM after the fixes, GP tests, ncu captures of
# the round-2 kernels, launch list of the bench step, then the bench (both arms)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gp_gpu.py tests/test_smc_gpu.py tests/test_summaries_gpu.py tests/test_bolfi_gpu.py -m gpu -q -k "not reference_bounds" 2>&1 | tail -6 > gpurun_out/r2f_pytest.log; tail -4 gpurun_out/r2f_pytest.log
TAG=v2 timeout 600 python scripts/bench_r2_kernels.py > gpurun_out/r2f_kernels.log 2>&1; grep -E 'nested|fused|colmoments|autocov12_4e5|meanvar_4e5' gpurun_out/r2f_kernels.log | cut -c1-250
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2f_bench_kernels.log 2>&1; grep -E "K10|K12|rank-1" gpurun_out/r2f_bench_kernels.log | cut -c1-300
for t in fused meanvar gm2; do
  case $t in
    fused)   k='regex:rowstream_kernel'; skip=1; cnt=1;;
    meanvar) k='regex:rowstream_kernel'; skip=1; cnt=4;;
    gm2)     k='regex:gm_pdf_kernel'; skip=1; cnt=1;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k "$k" -s $skip -c $cnt -f \
      -o gpurun_out/r2_prof_$t python scripts/prof_targets.py $t > gpurun_out/r2_ncu_$t.log 2>&1
  tail -1 gpurun_out/r2_ncu_$t.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r2_bench_under_ncu.log | cut -c1-200
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "ours rc=$?"; cut -c1-400 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'])
print(json.dumps(d.get('bolfi_config4'))[:1500])
print(json.dumps(d['smc_ma2'])[:800])
PY
