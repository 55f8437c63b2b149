#!/usr/bin/env python
"""argsort timing on the key patterns the multi-rank population gather produces."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elfi_b200 import ops  # noqa: E402


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


gen = torch.Generator(device='cuda').manual_seed(0)
n = 1_600_000
rand = torch.rand(n, dtype=torch.float64, device='cuda', generator=gen)
runs = torch.cat([torch.sort(rand[i * (n // 8):(i + 1) * (n // 8)])[0] for i in range(8)])
padded = runs.clone()
for i in range(8):
    padded[i * (n // 8) + 150_000:(i + 1) * (n // 8)] = float('inf')
strided = torch.rand(n, 3, dtype=torch.float64, device='cuda', generator=gen)
for name, keys in (('random', rand), ('8_sorted_runs', runs), ('8_sorted_runs_inf_padded', padded)):
    print(json.dumps({'name': 'argsort_1.6e6_' + name, 'ms': timeit(lambda: ops.argsort(keys))}), flush=True)
print(json.dumps({'name': 'strided_column_contiguous_copy_1.6e6x3',
                  'ms': timeit(lambda: strided[:, 0].contiguous())}))
perm = ops.argsort(rand)
print(json.dumps({'name': 'take_rows_1e6_of_1.6e6x3', 'ms': timeit(lambda: ops.take_rows(strided, perm[:1_000_000]))}))
