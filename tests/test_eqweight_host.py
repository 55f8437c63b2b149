"""CPU check of the closed-form equal-weight quantile position (elfi_b200/csrc/eqweight.h)
against the reference's own arithmetic (elfi/methods/utils.py:379-411 with weights=None):
np.cumsum of n copies of 1/n, last entry forced to 1, first k with cum[k] < alpha <= cum[k+1]."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def index_fn(tmp_path_factory):
    gxx = shutil.which('g++')
    if gxx is None:
        pytest.skip('g++ not available')
    so = str(tmp_path_factory.mktemp('eqw') / 'eqweight_harness.so')
    subprocess.check_call([gxx, '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-o',
                           so, os.path.join(HERE, 'harness', 'eqweight_harness.cpp')])
    lib = ctypes.CDLL(so)
    lib.harness_equal_weight_cum_index.restype = ctypes.c_int64
    lib.harness_equal_weight_cum_index.argtypes = [ctypes.c_int64, ctypes.c_double]
    return lib.harness_equal_weight_cum_index


def _reference_index(cum, alpha):
    """cum = [0, cumsum..., 1]; returns j = k + 1 of the reference's np.where(...)[0][0]."""
    return int(np.where(np.logical_and(cum[:-1] < alpha, alpha <= cum[1:]))[0][0]) + 1


def _cum(n):
    w = np.ones(n)
    w = w / np.sum(w)
    cum = np.insert(np.cumsum(w), 0, 0)
    cum[-1] = 1.0
    return cum


def _alphas(n, rs):
    ks = np.unique(np.clip(rs.randint(1, n + 1, 6), 1, n))
    out = [0.5, 0.25, 0.1, 0.01, 0.9, 0.99, 1.0, 1e-9, 1.0 / 3.0, 0.2, 0.3, 0.7]
    for k in ks:
        out += [k / n, np.nextafter(k / n, 0), np.nextafter(k / n, 1), k * (1.0 / n)]
    out += list(rs.rand(4))
    return [a for a in out if 0.0 < a <= 1.0]


def test_every_small_population(index_fn):
    rs = np.random.RandomState(0)
    for n in range(1, 1500):
        cum = _cum(n)
        for a in _alphas(n, rs):
            assert index_fn(n, a) == _reference_index(cum, a), (n, a)


@pytest.mark.parametrize('n', [4096, 10000, 65536, 99999, 100000, 1000000, 1048576, 1234567,
                               2000000, 3000001])
def test_large_populations_on_the_knife_edges(index_fn, n):
    rs = np.random.RandomState(n)
    cum = _cum(n)
    # the cumulative weights themselves and their neighbours are the hardest alphas
    picks = np.unique(np.concatenate([rs.randint(1, n, 40), [1, 2, n // 2, n - 1, n // 4, n // 10]]))
    alphas = _alphas(n, rs)
    for k in picks:
        alphas += [cum[k], np.nextafter(cum[k], 0), np.nextafter(cum[k], 1)]
    for a in alphas:
        if 0.0 < a <= 1.0:
            assert index_fn(n, a) == _reference_index(cum, a), (n, a)
