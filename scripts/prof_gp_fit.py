#!/usr/bin/env python
"""Three GP fits at n = 2000 (for an ncu launch list of elfi_b200_gp_fit_f64)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elfi_b200.bo import GPyRegression  # noqa: E402

rs = np.random.RandomState(0)
X = rs.uniform([-2, -1], [2, 1], (2000, 2))
y = np.log(0.05 + np.sum((X - 0.3) ** 2, axis=1)) + 0.1 * rs.randn(2000)
gp = GPyRegression(['t1', 't2'], bounds={'t1': (-2, 2), 't2': (-1, 1)}, incremental=False)
gp.update(X, y)
for _ in range(2):
    gp._fit()
torch.cuda.synchronize()
print('done')
