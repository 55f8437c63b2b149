#!/usr/bin/env python
"""torch.profiler view of one multi-rank SMC run (rank 0): runtime API calls (cudaMalloc, syncs),
NCCL kernels and our kernels around the population gather."""
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
local = int(os.environ.get('LOCAL_RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import elfi_b200 as elfi  # noqa: E402
from elfi_b200.examples import ma2  # noqa: E402

m = ma2.get_device_model(seed_obs=4)


def run(pops):
    smc = elfi.SMC(m['d'], batch_size=125000, seed=1, device_proposal=ma2.DeviceProposal)
    res = smc.sample(1_000_000, quantiles=[0.5] * pops, bar=False)
    torch.cuda.synchronize()
    return res


run(2)
run(2)
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(3)
if int(os.environ.get('RANK', '0')) == 0:
    print(prof.key_averages().table(sort_by='self_cuda_time_total', row_limit=22, max_name_column_width=60))
    print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=22, max_name_column_width=60))
if world > 1:
    dist.destroy_process_group()
