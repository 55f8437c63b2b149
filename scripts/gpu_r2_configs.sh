#!/bin/bash
# round 2: BASELINE configs 3 and 5 at scale on N ranks (N = 1 or 8), throughput mode.
#   config 3: Gaussian model, SMC 5 populations x 1e6 particles (quantiles) and AdaptiveThresholdSMC
#   config 5: g-and-k, AdaptiveDistanceSMC, 256 order statistics, population 2e6 at quantile 0.5
#             = 4e6 particles kept per generation (5e5 per GPU on 8 ranks), 3 rounds
mkdir -p gpurun_out
N=${1:-1}
if [ $N -eq 1 ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29641"; fi
export ELFI_B200_TIMING=${TIMING:-0}
timeout 900 $RUN scripts/bench_smc.py --model gauss --population 1000000 --batch 125000 --pops 5 > gpurun_out/r2_cfg3_gauss_smc_g$N.json 2> gpurun_out/r2_cfg3_gauss_smc_g$N.err; echo "gauss smc rc=$?"; cut -c1-900 gpurun_out/r2_cfg3_gauss_smc_g$N.json
timeout 900 $RUN scripts/bench_smc.py --model gauss --adaptive-threshold --population 1000000 --batch 125000 --pops 5 > gpurun_out/r2_cfg3_gauss_adathr_g$N.json 2> gpurun_out/r2_cfg3_gauss_adathr_g$N.err; echo "gauss adaptive-threshold rc=$?"; cut -c1-900 gpurun_out/r2_cfg3_gauss_adathr_g$N.json; tail -2 gpurun_out/r2_cfg3_gauss_adathr_g$N.err
timeout 1200 $RUN scripts/bench_smc.py --model gnk --n-obs 256 --population 2000000 --batch 500000 --pops 3 --quantile 0.5 > gpurun_out/r2_cfg5_gnk_g$N.json 2> gpurun_out/r2_cfg5_gnk_g$N.err; echo "gnk rc=$?"; cut -c1-1200 gpurun_out/r2_cfg5_gnk_g$N.json; tail -3 gpurun_out/r2_cfg5_gnk_g$N.err
