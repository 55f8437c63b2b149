import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session', autouse=True)
def _built():
    """Make sure the CUDA library and the oracle are built (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


@pytest.fixture
def golden():
    return load_golden


@pytest.fixture
def cpu_double(monkeypatch):
    """Host-logic tests without a GPU: the C ABI is replaced by tests/abi_double.py (NumPy + the
    oracle on host pointers) and device allocations by CPU tensors.  Test-only; see that module."""
    import abi_double
    abi_double.install(monkeypatch)
    return abi_double
