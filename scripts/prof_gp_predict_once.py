#!/usr/bin/env python
"""One GP fit (n = 2000) and one grid prediction (m = 1e5) -- the target of the ncu capture of the
triangular DMMA GEMM (`ncu --kernel-name-base demangled -k regex:'gemm_nt_dmma_kernel<.*true>'`)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from elfi_b200 import device as dev  # noqa: E402
from elfi_b200.bo import GPyRegression  # noqa: E402

rs = np.random.RandomState(0)
Xe = rs.uniform([-2, -1], [2, 1], (2000, 2))
ye = np.log(0.05 + np.sum((Xe - 0.3) ** 2, axis=1)) + 0.1 * rs.randn(2000)
gp = GPyRegression(['t1', 't2'], bounds={'t1': (-2, 2), 't2': (-1, 1)})
gp.update(Xe, ye)
g1, g2 = np.meshgrid(np.linspace(-2, 2, 400), np.linspace(-1, 1, 250))
grid = dev.to_device(np.column_stack([g1.ravel(), g2.ravel()]))
out = gp.predict_device(grid, noiseless=True, beta=20.0)
torch.cuda.synchronize()
print('ok', float(out[0][0]))
