#!/bin/bash
# gpurun --gpus N helper: multi-GPU sampler parity + bench scaling at 1..N GPUs
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
   --master-port 29611 tests/mgpu_check.py > gpurun_out/mgpu_check.log 2>&1
tail -5 gpurun_out/mgpu_check.log
for g in 1 2 4 8; do
  if [ $g -le $N ]; then
    if [ $g -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_g$g.json 2> gpurun_out/bench_g$g.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 \
        --master-port 2962$g bench.py --gpus $g --steps 20 --warmup 3 > gpurun_out/bench_g$g.json 2> gpurun_out/bench_g$g.err
    fi
    tail -1 gpurun_out/bench_g$g.json | cut -c1-400
  fi
done
