#!/bin/bash
# compact-code GP panel / block-inverse kernels: parity, then fit timing and its launch list
mkdir -p gpurun_out
timeout 300 python scripts/debug_gp_fit.py > gpurun_out/r2q_debug_gp_fit.log 2>&1; tail -30 gpurun_out/r2q_debug_gp_fit.log | cut -c1-200
timeout 1200 python -m pytest tests/test_gp_gpu.py tests/test_bolfi_gpu.py -x -q > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -4 gpurun_out/r2q_pytest.log
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2q_bench_kernels.log 2>&1
grep -i "GP \|gp_\|LCBSC\|rowsort" gpurun_out/r2q_bench_kernels.log | cut -c1-260
cp gpurun_out/kernels.json gpurun_out/r2q_kernels.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_gp_fit_launches.csv python scripts/prof_gp_fit.py > gpurun_out/r2q_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = []
with open('gpurun_out/r2q_gp_fit_launches.csv') as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
for row in r:
    name = row['Kernel Name'].split('(')[0]
    v = float(row['Metric Value'].replace(',', ''))
    if row.get('Metric Unit') == 'ns': v /= 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, t) in agg.items():
    print('%-60s launches %4d  total %9.1f us  avg %7.2f us' % (k[:60], c, t, t / c))
PY
