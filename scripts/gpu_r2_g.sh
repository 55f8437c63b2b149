#!/bin/bash
# round 2, call G (N GPUs): NCCL parity (mgpu_check), bench at N ranks, configs 3 and 5 at N ranks
N=${1:-8}
bash scripts/gpu_r2_b.sh $N
bash scripts/gpu_r2_configs.sh $N
