#!/bin/bash
# round-2 first call (2 GPUs): box topology + SMC phase timings at 1 and 2 ranks (baseline for the
# per-generation overhead work)
mkdir -p gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; numactl -H 2>/dev/null | head -20; nvidia-smi topo -m; free -g | head -2; } > gpurun_out/r2_box.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
ELFI_B200_TIMING=1 timeout 300 python scripts/bench_smc.py --population 1000000 --batch 125000 --pops 3 > gpurun_out/r2_diag_smc_g1.json 2> gpurun_out/r2_diag_smc_g1.err
cut -c1-1500 gpurun_out/r2_diag_smc_g1.json
ELFI_B200_TIMING=1 timeout 300 $TR --nproc-per-node 2 --master-port 29631 scripts/bench_smc.py --population 1000000 --batch 125000 --pops 3 > gpurun_out/r2_diag_smc_g2.json 2> gpurun_out/r2_diag_smc_g2.err
cut -c1-1500 gpurun_out/r2_diag_smc_g2.json
timeout 300 $TR --nproc-per-node 2 --master-port 29632 scripts/bench_smc.py --population 1000000 --batch 125000 --pops 3 > gpurun_out/r2_diag_smc_g2_notiming.json 2>> gpurun_out/r2_diag_smc_g2.err
cut -c1-600 gpurun_out/r2_diag_smc_g2_notiming.json
tail -3 gpurun_out/r2_diag_smc_g2.err
