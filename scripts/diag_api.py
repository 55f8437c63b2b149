import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elfi_b200 as elfi
from elfi_b200.examples import ma2
from elfi_b200.samplers import PHASES
m = ma2.get_device_model(seed_obs=4)
B = 1_000_000
for rep in range(3):
    PHASES.tot.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = elfi.Rejection(m['d'], batch_size=B, seed=2 + rep, distributed=False).sample(8 * B // 100, n_sim=8 * B, bar=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('rep', rep, 'ms/batch', dt / 8 * 1e3, PHASES.report(), flush=True)
