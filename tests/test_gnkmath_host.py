"""CPU check of the g-and-k quantile function the device simulator evaluates
(elfi_b200/csrc/gnkmath.cuh vs the expression of elfi/examples/gnk.py:60-66)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def harness(tmp_path_factory):
    gxx = shutil.which('g++')
    if gxx is None:
        pytest.skip('g++ not available')
    so = str(tmp_path_factory.mktemp('gnk') / 'gnk_harness.so')
    subprocess.check_call([gxx, '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-o',
                           so, os.path.join(HERE, 'harness', 'gnk_harness.cpp')])
    return ctypes.CDLL(so)


@pytest.mark.parametrize('prm', [(3.0, 1.0, 2.0, 0.5), (0.0, 10.0, 0.0, 0.0), (7.5, 0.3, 9.9, 4.2),
                                 (1.0, 2.0, 0.1, 10.0)])
def test_quantile_function_matches_reference_expression(harness, prm):
    rs = np.random.RandomState(0)
    z = np.concatenate([rs.randn(5000), [0.0, -0.0, 6.5, -6.5, 1e-300]])
    out = np.empty_like(z)
    p = np.asarray(prm, dtype=np.float64)
    harness.harness_gnk_quantile(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(0.8),
                                 z.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(z)),
                                 out.ctypes.data_as(ctypes.c_void_p))
    A, B, g, k = prm
    c = 0.8
    ref = A + B * (1 + c * ((1 - np.exp(-g * z)) / (1 + np.exp(-g * z)))) * (1 + z**2)**k * z
    assert np.allclose(out, ref, rtol=1e-13, atol=0.0)
