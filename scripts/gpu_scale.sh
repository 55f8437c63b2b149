#!/bin/bash
# gpurun --gpus N: multi-GPU parity check + scaling of the headline bench and of SMC-ABC MA2
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29611 tests/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1
grep -E "MGPU_OK|Error" gpurun_out/mgpu_check_$N.log | head -3
for g in 1 2 4 8; do
  if [ $g -le $N ]; then
    if [ $g -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_bench_g$g.json 2> gpurun_out/scale_bench_g$g.err
      timeout 600 python scripts/bench_smc.py --population 1000000 --batch 1000000 --pops 3 > gpurun_out/scale_smc_g$g.json 2> gpurun_out/scale_smc_g$g.err
    else
      timeout 600 $TR --nproc-per-node $g --master-port 2962$g bench.py --gpus $g --steps 20 --warmup 3 > gpurun_out/scale_bench_g$g.json 2> gpurun_out/scale_bench_g$g.err
      timeout 600 $TR --nproc-per-node $g --master-port 2963$g scripts/bench_smc.py --population 1000000 --batch $((1000000 / g)) --pops 3 > gpurun_out/scale_smc_g$g.json 2> gpurun_out/scale_smc_g$g.err
    fi
    tail -1 gpurun_out/scale_bench_g$g.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench g=%d value=%.4g ms=%.4f e2e=%.4g frac=%.3f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac']))"
    tail -1 gpurun_out/scale_smc_g$g.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('smc g=%d s=%.3f acc/s=%.4g pairs/s=%.4g'%(d['n_gpus'],d['seconds'],d['accepted_particles_per_s'],d['pair_terms_per_s']))"
  fi
done
