#!/usr/bin/env python
"""torch.profiler view of throughput-mode Rejection (device MA2 model, batch 1e6, 8 batches):
which kernels and which host calls make up a batch."""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import elfi_b200 as elfi  # noqa: E402
from elfi_b200.examples import ma2  # noqa: E402

B = 1_000_000
m = ma2.get_device_model(seed_obs=4)


def run(seed):
    res = elfi.Rejection(m['d'], batch_size=B, seed=seed, distributed=False).sample(
        8 * B // 100, n_sim=8 * B, bar=False)
    float(res.sample_means['t1'])
    torch.cuda.synchronize()


run(1)
for s in (2, 3):
    t0 = time.perf_counter()
    run(s)
    print('wall ms per batch %.3f' % ((time.perf_counter() - t0) / 8 * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(4)
print(prof.key_averages().table(sort_by='self_cuda_time_total', row_limit=16, max_name_column_width=70))
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=14, max_name_column_width=60))
