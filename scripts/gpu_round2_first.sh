#!/bin/bash
# First GPU call of round 2 (one B200): everything that was written after round 1's GPU budget
# was spent gets its first device run and its first timing here.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# 1. parity: full GPU suite (new: g-and-k throughput mode, pools, BOLFI sampling path, closed-form
#    equal-weight quantile, signed-zero summaries)
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.log
cat gpurun_out/r2_pytest_gpu.log
# 2. timings: summary kernels (leaf + tree), per-kernel table, weighted quantile with equal weights
timeout 300 python scripts/gpu_leaf_check.py > gpurun_out/r2_leaf_check.log 2>&1; tail -2 gpurun_out/r2_leaf_check.log | cut -c1-400
# same with the term-wise tree accumulator (treesum.cuh) switched on: parity, then the 4e5 x 256 timing
ELFI_B200_SUMM_TERMWISE=1 timeout 300 python scripts/gpu_leaf_check.py > gpurun_out/r2_leaf_check_termwise.log 2>&1; tail -2 gpurun_out/r2_leaf_check_termwise.log | cut -c1-400
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2_bench_kernels.log 2>&1; tail -3 gpurun_out/r2_bench_kernels.log | cut -c1-300
# 3. SMC throughput mode, 1 GPU: MA2 (round-0 quantile now closed form), Gaussian (config #3),
#    g-and-k with AdaptiveDistanceSMC (config #5)
for model in ma2 gauss; do
  ELFI_B200_TIMING=1 timeout 600 python scripts/bench_smc.py --model $model --population 1000000 --pops 3 \
      > gpurun_out/r2_smc_${model}.json 2> gpurun_out/r2_smc_${model}.err; cut -c1-400 gpurun_out/r2_smc_${model}.json
done
# config #3 proper: adaptive threshold (KLIEP) on the Gaussian model, 5 populations x 1e6 particles
ELFI_B200_TIMING=1 timeout 900 python scripts/bench_smc.py --model gauss --adaptive-threshold --population 1000000 --pops 5 \
    > gpurun_out/r2_smc_gauss_adaptive.json 2> gpurun_out/r2_smc_gauss_adaptive.err; cut -c1-400 gpurun_out/r2_smc_gauss_adaptive.json; tail -2 gpurun_out/r2_smc_gauss_adaptive.err
ELFI_B200_TIMING=1 timeout 600 python scripts/bench_smc.py --model gnk --population 100000 --pops 3 --n-obs 256 \
    > gpurun_out/r2_smc_gnk.json 2> gpurun_out/r2_smc_gnk.err; cut -c1-400 gpurun_out/r2_smc_gnk.json; tail -2 gpurun_out/r2_smc_gnk.err
# 4. headline bench, both arms
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cut -c1-500 gpurun_out/r2_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench.err; cut -c1-300 gpurun_out/r2_bench_ref.json
