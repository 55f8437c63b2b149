#!/usr/bin/env python
"""Launch each hot kernel a few times at its BASELINE shape (for ncu captures)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elfi_b200 import ops  # noqa: E402

gen = torch.Generator(device='cuda').manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
reps = 3
if which in ('all', 'dist'):
    S = torch.randn(1_000_000, 128, dtype=torch.float64, device='cuda', generator=gen)
    obs = torch.randn(128, dtype=torch.float64, device='cuda', generator=gen)
    for _ in range(reps):
        ops.dist_euclid(S, obs, thresholds=13.3)
    del S
if which in ('all', 'nested'):
    S5 = torch.randn(500_000, 256, dtype=torch.float64, device='cuda', generator=gen)
    obs5 = torch.randn(256, dtype=torch.float64, device='cuda', generator=gen)
    W5 = torch.rand(5, 256, dtype=torch.float64, device='cuda', generator=gen) + 0.5
    for _ in range(reps):
        ops.dist_euclid(S5, obs5, w=W5)
    del S5
if which in ('all', 'summ'):
    X = torch.randn(1_000_000, 100, dtype=torch.float64, device='cuda', generator=gen)
    for _ in range(reps):
        ops.autocov(X, lags=(1, 2))
    Y = torch.randn(1_000_000, 50, dtype=torch.float64, device='cuda', generator=gen)
    for _ in range(reps):
        ops.meanvar(Y)
    del X, Y
if which in ('all', 'gm'):
    x = torch.randn(100_000, 2, dtype=torch.float64, device='cuda', generator=gen)
    m = torch.randn(50_000, 2, dtype=torch.float64, device='cuda', generator=gen)
    w = torch.rand(50_000, dtype=torch.float64, device='cuda', generator=gen)
    for _ in range(2):
        ops.gm_logpdf(x, m, np.diag([0.05, 0.02]), w)
if which in ('all', 'gp'):
    from elfi_b200.bo import GPyRegression
    rs = np.random.RandomState(0)
    Xe = rs.uniform([-2, -1], [2, 1], (2000, 2))
    ye = np.log(0.05 + np.sum((Xe - 0.3) ** 2, axis=1)) + 0.1 * rs.randn(2000)
    gp = GPyRegression(['t1', 't2'], bounds={'t1': (-2, 2), 't2': (-1, 1)})
    gp.update(Xe, ye)
    grid = rs.uniform([-2, -1], [2, 1], (16384, 2))
    gp.predict_device(grid, noiseless=True, beta=20.0)
if which in ('fused',):     # round 2: nested distances (K = 5) + column moments from one read of S
    S5 = torch.randn(500_000, 256, dtype=torch.float64, device='cuda', generator=gen)
    obs5 = torch.randn(256, dtype=torch.float64, device='cuda', generator=gen)
    W5 = torch.rand(5, 256, dtype=torch.float64, device='cuda', generator=gen) + 0.5
    for _ in range(reps):
        ops.dist_euclid(S5, obs5, w=W5, thresholds=np.full(5, 1e9), sync=False, moments=True)
    del S5
if which in ('meanvar',):   # round 2: single-sweep mean / variance (Gaussian model summaries)
    Y = torch.randn(1_000_000, 50, dtype=torch.float64, device='cuda', generator=gen)
    for _ in range(reps):
        ops.meanvar(Y)
    Z = torch.randn(400_000, 256, dtype=torch.float64, device='cuda', generator=gen)
    for _ in range(reps):
        ops.autocov(Z, lags=(1, 2))
    del Y, Z
if which in ('gm2',):       # round 2: mixture density at one rank's shard of the 1e6 x 1e6 problem / 8
    x = torch.randn(125_000, 2, dtype=torch.float64, device='cuda', generator=gen) * 0.3
    m = torch.randn(125_000, 2, dtype=torch.float64, device='cuda', generator=gen) * 0.3
    w = torch.rand(125_000, dtype=torch.float64, device='cuda', generator=gen) + 0.1
    for _ in range(2):
        ops.gm_logpdf(x, m, np.array([[0.02, 0.004], [0.004, 0.01]]), w, validate=False)
torch.cuda.synchronize()
print('done')
