"""Plugin boundary #2: elfi_b200.client.Client driven by the UNMODIFIED
reference's BatchHandler / samplers.  Needs /root/reference, so it only runs in the build
container (skipped on the GPU box); the node operations here are the reference's host ones."""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/elfi'),
                                reason='reference checkout not present')


def test_client_runs_reference_rejection():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    'oracle'))
    from ref_shim import import_reference
    elfi = import_reference()
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
    c = Client(devices=[0, 1])
    a = c.apply(lambda x, y=1: x + y, 2, y=3)
    b = c.apply(lambda: 7)
    assert c.is_ready(a) and c.num_cores == 2
    c.remove_task(b)
    assert c.get_result(a) == 5
    with pytest.raises(KeyError):
        c.get_result(b)
    assert c.apply_sync(lambda: 'ok') == 'ok'
    c.reset()
    assert not c.tasks
