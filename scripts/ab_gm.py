import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elfi_b200 import ops
gen = torch.Generator(device='cuda').manual_seed(0)
M = 1_000_000
mns = torch.randn(M, 2, dtype=torch.float64, device='cuda', generator=gen)
wm = torch.rand(M, dtype=torch.float64, device='cuda', generator=gen)
cov = np.diag([0.05, 0.02])
for N in (125_000, 1_000_000):
    x = torch.randn(N, 2, dtype=torch.float64, device='cuda', generator=gen)
    for rep in range(2):
        for waves in ('1', '4', '16', '64'):
            os.environ['ELFI_B200_GM_WAVES'] = waves
            ops.gm_logpdf(x[:1000], mns[:1000], cov, wm[:1000])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.gm_logpdf(x, mns, cov, wm); b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            print('N=%d M=%d waves=%s ms=%.1f pairs/s=%.3e' % (N, M, waves, ms, N * M / ms * 1e3), flush=True)
