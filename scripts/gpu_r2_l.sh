#!/bin/bash
# round 2, call L (1 GPU): GP fit after the register factorisation + look-ahead: parity, timing, launch list
mkdir -p gpurun_out
timeout 60 python scripts/debug_gp_fit.py 2>&1 | grep "max err"
timeout 600 python -m pytest tests/test_gp_gpu.py tests/test_bolfi_gpu.py -m gpu -q 2>&1 | tail -6
timeout 300 python scripts/bench_kernels.py > gpurun_out/r2l_bench_kernels.log 2>&1; grep -E "K10|K12|rank-1" gpurun_out/r2l_bench_kernels.log | cut -c1-200
ELFI_B200_GP_LOOKAHEAD=0 timeout 300 python scripts/bench_kernels.py > gpurun_out/r2l_bench_kernels_nolook.log 2>&1; grep -E "K10" gpurun_out/r2l_bench_kernels_nolook.log | cut -c1-200
timeout 120 python scripts/bench_sort_inputs.py 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_gp_fit_launches.csv python scripts/prof_gp_fit.py > gpurun_out/r2l_ncu.log 2>&1; tail -1 gpurun_out/r2l_ncu.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_gp_fit_launches.csv')) if len(r)>14 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][-48:]
    agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=float(r[14].replace(',',''))
for k,(n,t) in agg.items(): print('%5d %10.1f us total %8.1f us/launch  %s'%(n,t/1e3,t/n/1e3,k))
PY
