"""CPU check of the single-leaf summary arithmetic (elfi_b200/csrc/leafsum.cuh).

The structs in that header are the per-lane state of the row-stream summary kernels for rows of
<= 128 terms; they also compile for the host.  tests/harness/leaf_harness.cpp feeds them rows 16
columns at a time exactly like the kernel does (NaN beyond the row end) and this test compares
the results with NumPy bit for bit -- the same comparison tests/test_summaries_gpu.py makes for
the CUDA build (reference: elfi/examples/ma2.py:40-59, gauss.py:142-173).
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PAIRS = [(1, 2), (1, -1), (2, -1), (3, -1), (4, -1)]


@pytest.fixture(scope='module')
def harness(tmp_path_factory):
    gxx = shutil.which('g++')
    if gxx is None:
        pytest.skip('g++ not available')
    so = str(tmp_path_factory.mktemp('leaf') / 'leaf_harness.so')
    subprocess.check_call([gxx, '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-o',
                           so, os.path.join(HERE, 'harness', 'leaf_harness.cpp')])
    return ctypes.CDLL(so)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _autocov(lib, x, la, lb):
    out = np.empty((x.shape[0], 2))
    rc = lib.harness_autocov(_ptr(x), ctypes.c_int64(x.strides[0] // 8),
                             ctypes.c_int64(x.shape[0]), x.shape[1], la, lb, _ptr(out))
    assert rc == 0
    return out


def _meanvar(lib, x):
    out = np.empty((x.shape[0], 2))
    lib.harness_meanvar(_ptr(x), ctypes.c_int64(x.strides[0] // 8), ctypes.c_int64(x.shape[0]),
                        x.shape[1], _ptr(out))
    return out


def _same_bits(a, b):
    return np.array_equal(np.asarray(a).view(np.int64), np.asarray(b).view(np.int64))


def _rows(rs, B, n):
    x = rs.randn(B, n) * rs.choice([1e-3, 1.0, 1e6], size=(B, 1))
    x[0] = 0.0
    x[1] = -0.0
    x[2, ::2] = 0.0
    return x


def test_autocov_every_length(harness):
    rs = np.random.RandomState(11)
    for n in range(2, 134):
        x = _rows(rs, 48, n)
        for la, lb in PAIRS:
            if max(la, lb) >= n or n - la > 128:
                continue
            got = _autocov(harness, x, la, lb)
            assert _same_bits(got[:, 0], np.mean(x[:, la:] * x[:, :-la], axis=1)), (n, la)
            if lb > 0:
                assert _same_bits(got[:, 1], np.mean(x[:, lb:] * x[:, :-lb], axis=1)), (n, lb)


def test_meanvar_every_length(harness):
    rs = np.random.RandomState(12)
    for n in range(1, 129):
        x = _rows(rs, 48, n)
        x[3:] += 1.0
        got = _meanvar(harness, x)
        assert _same_bits(got[:, 0], np.mean(x, axis=1)), n
        assert _same_bits(got[:, 1], np.var(x, axis=1)), n


def test_meanvar_single_sweep_every_length(harness):
    """MeanVarRegs (row kept in registers, one sweep): every row length it serves, 1..64."""
    rs = np.random.RandomState(15)
    for n in range(1, 65):
        x = _rows(rs, 48, n)
        x[3:] += 1.0
        out = np.empty((x.shape[0], 2))
        rc = harness.harness_meanvar_regs(_ptr(x), ctypes.c_int64(x.strides[0] // 8),
                                          ctypes.c_int64(x.shape[0]), n, _ptr(out))
        assert rc == 0
        assert _same_bits(out[:, 0], np.mean(x, axis=1)), n
        assert _same_bits(out[:, 1], np.var(x, axis=1)), n
    from conftest import load_golden
    g = load_golden('gauss_generate')
    x = np.ascontiguousarray(g['gauss'])
    out = np.empty((x.shape[0], 2))
    harness.harness_meanvar_regs(_ptr(x), ctypes.c_int64(x.strides[0] // 8),
                                 ctypes.c_int64(x.shape[0]), x.shape[1], _ptr(out))
    assert _same_bits(out[:, 0], g['ss_mean']) and _same_bits(out[:, 1], g['ss_var'])


def test_strided_rows_and_golden(harness):
    """Row stride larger than the row length, and the reference's own MA2 / Gaussian batches."""
    from conftest import load_golden
    rs = np.random.RandomState(13)
    big = rs.randn(100, 160)
    x = big[:, :100]
    got = _autocov(harness, x, 1, 2)
    assert _same_bits(got[:, 0], np.mean(x[:, 1:] * x[:, :-1], axis=1))
    g = load_golden('ma2_generate')
    got = _autocov(harness, np.ascontiguousarray(g['MA2']), 1, 2)
    assert _same_bits(got[:, 0], g['S1']) and _same_bits(got[:, 1], g['S2'])
    g = load_golden('gauss_generate')
    got = _meanvar(harness, np.ascontiguousarray(g['gauss']))
    assert _same_bits(got[:, 0], g['ss_mean']) and _same_bits(got[:, 1], g['ss_var'])


# ------------------------------------------------------------------ TreeSum (treesum.cuh)
def _tree_call(lib, fn, x, *args):
    out = np.empty((x.shape[0], 2))
    rc = getattr(lib, fn)(_ptr(x), ctypes.c_int64(x.strides[0] // 8), ctypes.c_int64(x.shape[0]),
                          x.shape[1], *args, _ptr(out))
    assert rc == 0
    return out


def test_tree_accumulator_matches_numpy(harness):
    """Rows longer than one leaf: every length up to 300, the split points around powers of two,
    and the longest run a 6-level stack can hold (7688)."""
    rs = np.random.RandomState(14)
    lengths = [2, 7, 8, 9, 100, 127, 128] + list(range(129, 300)) + [383, 384, 385, 511, 512, 513, 1000, 1023, 1024, 1025, 1031,
                                     2047, 2048, 2049, 4099, 5000, 7679, 7680, 7681, 7687, 7688]
    for n in lengths:
        x = _rows(rs, 8, n)
        for la, lb in PAIRS:
            if max(la, lb) >= n:
                continue
            got = _tree_call(harness, 'harness_autocov_tree', x, la, lb)
            assert _same_bits(got[:, 0], np.mean(x[:, la:] * x[:, :-la], axis=1)), (n, la)
            if lb > 0:
                assert _same_bits(got[:, 1], np.mean(x[:, lb:] * x[:, :-lb], axis=1)), (n, lb)
        got = _tree_call(harness, 'harness_meanvar_tree', x)
        assert _same_bits(got[:, 0], np.mean(x, axis=1)), n
        assert _same_bits(got[:, 1], np.var(x, axis=1)), n


def test_stack_depth_bound():
    """max_terms() of PairwiseStream / TreeSum: a run of n terms needs as many stack levels as
    splits are open at once; right parts have up to n/2 + 7 terms, so 120 * 2^D + 8 terms is the
    most that D levels hold for every length below it (not 128 * 2^D)."""
    from functools import lru_cache

    @lru_cache(None)
    def levels(n):
        if n <= 128:
            return 0
        left = n // 2
        left -= left % 8
        return 1 + max(levels(left), levels(n - left))
    for depth in (1, 2, 6, 7, 8):
        bound = 120 * 2 ** depth + 8
        assert max(levels(m) for m in range(1, bound + 1)) == depth
        assert levels(bound + 1) == depth + 1
