#!/usr/bin/env python
"""CUDA-event timings of the fused device simulators (1e6 rows)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elfi_b200 import ops  # noqa: E402


def timeit(fn, per_batch=5, batches=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(batches):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(per_batch):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / per_batch)
    return float(np.median(ts))


B = 1_000_000
rs = np.random.RandomState(0)
t1 = torch.from_numpy(rs.uniform(-1, 1, B)).cuda()
t2 = torch.from_numpy(rs.uniform(0, 1, B)).cuda()
sg = torch.from_numpy(rs.uniform(0.5, 2, B)).cuda()
tag = os.environ.get('TAG', '')
print(tag, 'sim_ma2 1e6 x 100 fused summaries: %.3f ms' % timeit(lambda: ops.sim_ma2(t1, t2, 100, seed=1)))
for n in (50, 64):
    print(tag, 'sim_gauss 1e6 x %d fused summaries: %.3f ms' % (n, timeit(lambda: ops.sim_gauss(t1, sg, n, seed=1))))
print(tag, 'sim_gauss 1e6 x 50 data + summaries: %.3f ms' % timeit(lambda: ops.sim_gauss(t1, sg, 50, seed=1, want_data=True)))
