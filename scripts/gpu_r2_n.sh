#!/bin/bash
# round 2, call N (N GPUs): reference-plugin test (1 GPU), bench at N ranks
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_reference_plugin.py tests/test_merge_gpu.py -m gpu -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node $N --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2n_bench_g$N.json 2> gpurun_out/r2n_bench_g$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2n_bench_g$N.json').read().strip().splitlines()[-1])
print('value %.4g ms %.4f host %.4f e2e %.4g e2e_ms %.2f' % (d['value'], d['ms_per_step'], d['host_launch_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
s=d['smc_ma2']; print('smc', s['seconds'], s.get('mgpu_parity'), s['per_generation_ms'])
PY
tail -3 gpurun_out/r2n_bench_g$N.err | cut -c1-200
