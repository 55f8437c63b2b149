#!/bin/bash
# round 2, call B (N GPUs): NCCL sampler parity (mgpu_check), bench at N ranks
mkdir -p gpurun_out profiles
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29611 tests/mgpu_check.py > gpurun_out/r2_mgpu_check_w$N.log 2>&1
grep -E "MGPU_OK|Error|assert" gpurun_out/r2_mgpu_check_w$N.log | head -5
timeout 900 $TR --nproc-per-node $N --master-port 29621 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_g$N.json 2> gpurun_out/r2_bench_g$N.err
echo "bench rc=$?"; grep MGPU_OK gpurun_out/r2_bench_g$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_g$N.json').read().strip().splitlines()[-1])
print('value %.4g ms %.4f e2e %.4g e2e_ms %.2f pcie_frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pcie_frac']))
print(json.dumps(d['smc_ma2'])[:3000])
PY
tail -5 gpurun_out/r2_bench_g$N.err
