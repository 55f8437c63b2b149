#!/usr/bin/env python
"""Host-side profile (cProfile) of one SMC-ABC MA2 throughput-mode run: where the Python thread
spends its time (launch overhead, synchronising reads, graph execution) next to the kernels."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import elfi_b200 as elfi  # noqa: E402
from elfi_b200.examples import ma2  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 125_000
pops = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = ma2.get_device_model(seed_obs=4)


def run(n, batch, pops):
    smc = elfi.SMC(m['d'], batch_size=batch, seed=1, device_proposal=ma2.DeviceProposal)
    res = smc.sample(n, quantiles=[0.5] * pops, bar=False)
    _ = res.outputs['d'], res.weights
    torch.cuda.synchronize()
    return res


run(n, batch, 2)
t0 = time.perf_counter()
run(n, batch, pops)
print('plain run: %.4f s' % (time.perf_counter() - t0))
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
run(n, batch, pops)
dt = time.perf_counter() - t0
pr.disable()
print('profiled run: %.4f s' % dt)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25)
print(s.getvalue()[:5000])
