#!/bin/bash
# round 2, call AC (1 GPU): the driver's round-end sequence -- full GPU suite, smoke, both bench arms
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2ac_pytest_gpu.log; cat gpurun_out/r2ac_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2ac_bench_ref.json 2> gpurun_out/r2ac_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; echo "ours rc=$?"
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r2ac_bench_ref.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/r2ac_bench.json').read().strip().splitlines()[-1])
print('ref value %.4g ms %.2f' % (r['value'], r['ms_per_step']), 'same_config', r['config']==d['config'], 'crc', r['parity_check']==d['parity_check'])
print({k: d[k] for k in ('value','ms_per_step','host_launch_ms_per_step')}, 'e2e', d['e2e']['value'], 'ratio e2e', d['e2e']['value']/r['e2e']['value'], 'frac', d['roofline']['frac'])
print('smc', d['smc_ma2']['seconds'], d['smc_ma2']['per_generation_ms'])
print('api', d['api_throughput_mode'].get('vs_reference_arm'))
print('bolfi', {k: d['bolfi_config4'].get(k) for k in ('fit_ms','rank5_update_ms','grid_predict_lcbsc_ms','speedup_fit','speedup_grid')})
PY
tail -3 gpurun_out/r2ac_bench.err
timeout 300 python scripts/time_gp_predict.py --chunks > gpurun_out/r2ac_gp_predict_chunks.jsonl 2> gpurun_out/r2ac_gp_predict_chunks.err; cut -c1-400 gpurun_out/r2ac_gp_predict_chunks.jsonl; tail -3 gpurun_out/r2ac_gp_predict_chunks.err
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_nt_dmma_kernel<.*true>' -c 1 -f -o gpurun_out/r2ac_gemm_tri python scripts/prof_gp_predict_once.py > gpurun_out/r2ac_ncu_gemm.log 2>&1; tail -2 gpurun_out/r2ac_ncu_gemm.log
