"""Plugin boundaries of SURVEY.md section 8b, driven by the UNMODIFIED reference (from
/root/reference in the build container, from the mirror baseline/_ref on the GPU box):
  * #2 the client: elfi_b200.client.Client under the reference's BatchHandler / samplers
    (CPU: the node operations are the reference's host ones);
  * #1 the operators: the reference's own Rejection on a model whose Distance and Summary
    operations are the ctypes stubs of integration/elfi_b200_ops.py,
    i.e. the CUDA kernels inside an unmodified ELFI -- GPU test, bit-equal to the golden."""
import os
import sys
from functools import partial

import numpy as np
import pytest

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(),
                                reason='neither /root/reference nor baseline/_ref is present')


@pytest.mark.gpu
def test_reference_rejection_with_device_operators():
    """elfi.Rejection (reference code, reference client loop) + device Distance / Summary."""
    elfi = ref_shim.import_reference()
    from elfi.examples import ma2
    sys.path.insert(0, os.path.join(ROOT, 'integration'))
    import elfi_b200_ops as b200
    import importlib
    import elfi_b200.client as bc
    importlib.reload(bc)
    g = load_golden('ma2_rejection_quantile')
    old = elfi.client.get_client()
    try:
        elfi.set_client(bc.Client(devices=[0]))
        y = ma2.MA2(.6, .2, random_state=np.random.RandomState(4))
        m = elfi.ElfiModel()
        elfi.Prior(ma2.CustomPrior1, 2, model=m, name='t1')
        elfi.Prior(ma2.CustomPrior2, m['t1'], 1, name='t2')
        elfi.Simulator(partial(ma2.MA2, n_obs=100), m['t1'], m['t2'], observed=y, name='MA2')
        elfi.Summary(b200.device_autocov, m['MA2'], name='S1')
        elfi.Summary(partial(b200.device_autocov, lag=2), m['MA2'], name='S2')
        elfi.Distance(b200.device_cdist_euclidean, m['S1'], m['S2'], name='d')
        res = elfi.Rejection(m['d'], batch_size=1000, seed=123).sample(100, quantile=0.01,
                                                                       bar=False)
    finally:
        elfi.set_client(old)
    assert res.n_sim == int(g['n_sim']) and res.threshold == float(g['threshold'])
    assert np.array_equal(res.discrepancies, g['out_d'])
    assert np.array_equal(res.samples['t1'], g['out_t1'])
    assert np.array_equal(res.samples['t2'], g['out_t2'])
    # the reference's own SMC (proposals, importance weights, weighted quantile: all reference
    # host code) on the same device-operator model: every population equals the golden of the
    # all-host reference run bit for bit
    gs = load_golden('ma2_smc_thresholds')
    try:
        elfi.set_client(bc.Client(devices=[0]))
        smc = elfi.SMC(m['d'], batch_size=1000, seed=20).sample(150, thresholds=[.6, .3, .15],
                                                                bar=False)
    finally:
        elfi.set_client(old)
    assert len(smc.populations) == int(gs['n_pops']) and smc.n_sim == int(gs['n_sim'])
    for i, pop in enumerate(smc.populations):
        assert np.array_equal(pop.discrepancies, gs['pop{}_out_d'.format(i)])
        assert np.array_equal(pop.samples['t1'], gs['pop{}_out_t1'.format(i)])
        assert np.array_equal(pop.weights, gs['pop{}_weights'.format(i)])
    # the stubs alone against SciPy / NumPy
    from scipy.spatial.distance import cdist
    rs = np.random.RandomState(5)
    S, obs, w = rs.randn(3000, 128), rs.randn(1, 128), rs.rand(128)
    assert np.array_equal(b200.device_cdist_euclidean(S, obs), cdist(S, obs, 'euclidean'))
    assert np.array_equal(b200.device_cdist_euclidean(S, obs, w=w), cdist(S, obs, 'euclidean', w=w))
    x = rs.randn(500, 100)
    assert np.array_equal(b200.device_autocov(x, 2), np.mean(x[:, 2:] * x[:, :-2], axis=1))


def test_client_runs_reference_rejection():
    elfi = ref_shim.import_reference()
    from elfi.examples import ma2
    import importlib
    import elfi_b200.client as bc
    importlib.reload(bc)                       # pick up elfi.client.ClientBase as the base class
    assert issubclass(bc.Client, elfi.client.ClientBase)
    old = elfi.client.get_client()
    try:
        elfi.set_client(bc.Client(devices=[0]))
        assert elfi.client.get_client().num_cores == 1
        m = ma2.get_model(seed_obs=4)
        res = elfi.Rejection(m['d'], batch_size=1000, seed=123).sample(100, quantile=0.01,
                                                                       bar=False)
    finally:
        elfi.set_client(old)
    g = load_golden('ma2_rejection_quantile')
    assert res.n_sim == int(g['n_sim'])
    assert np.array_equal(res.discrepancies, g['out_d'])


def test_client_task_api():
    from elfi_b200.client import Client
    c = Client(devices=[0])
    a = c.apply(lambda x, y=1: x + y, 2, y=3)
    b = c.apply(lambda: 7)
    assert c.is_ready(a) and c.num_cores == 1
    c.remove_task(b)
    assert c.get_result(a) == 5
    with pytest.raises(KeyError):
        c.get_result(b)
    assert c.apply_sync(lambda: 'ok') == 'ok'
    c.reset()
    assert not c.tasks


@pytest.mark.gpu
def test_client_spreads_batches_over_devices():
    """Two worker threads (both on GPU 0 here, any two GPUs on a bigger box): tasks land on
    devices round-robin, run concurrently and come back by id."""
    import threading
    import torch
    from elfi_b200.client import Client
    n = min(2, torch.cuda.device_count())
    c = Client(devices=list(range(n)) if n == 2 else [0, 0])
    assert c.num_cores == 2
    seen = []

    def task(i):
        seen.append((i, torch.cuda.current_device(), threading.get_ident()))
        return torch.full((4,), float(i), device='cuda').sum().item()
    ids = [c.apply(task, i) for i in range(6)]
    assert [c.get_result(t) for t in ids] == [4.0 * i for i in range(6)]
    assert len({ident for _, _, ident in seen}) == 2
    for i, device, _ in seen:
        assert device == c.devices[i % 2]
