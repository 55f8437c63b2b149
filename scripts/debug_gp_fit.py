#!/usr/bin/env python
"""Compare the device Cholesky factor of elfi_b200_gp_fit_f64 with NumPy, block by block."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import elfi_oracle as o  # noqa: E402
from elfi_b200.bo import JITTER, GPyRegression  # noqa: E402

for n in (100, 300, 700):
    rs = np.random.RandomState(n)
    X = rs.uniform([-2, -1], [2, 1], (n, 2))
    y = np.log(0.05 + np.sum((X - 0.3) ** 2, axis=1)) + 0.1 * rs.randn(n)
    gp = GPyRegression(['t1', 't2'], bounds={'t1': (-2, 2), 't2': (-1, 1)}, incremental=False)
    gp.update(X, y[:, None])
    h = gp.hyperparameters
    f = gp._factor
    L = f['L'].cpu().numpy()[:n, :n]
    Lr, alpha = o.gp_fit(X, y, h['kernel_var'], h['lengthscale'], h['bias_var'], h['noise_var'], jitter=JITTER)
    err = np.abs(np.tril(L) - Lr)
    nb = (n + 63) // 64
    blocks = np.array([[err[i * 64:(i + 1) * 64, j * 64:(j + 1) * 64].max() if j <= i else 0
                        for j in range(nb)] for i in range(nb)])
    print('n', n, 'max err', err.max(), 'alpha err', np.abs(f['alpha'].cpu().numpy() - alpha.ravel()).max())
    with np.printoptions(precision=1, linewidth=200):
        print(blocks)
