"""GPU parity: summary kernels vs the oracle (== NumPy) and the reference goldens, bit-exact."""
import numpy as np
import pytest

import elfi_oracle as o
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,n', [(1000, 100), (333, 16), (64, 17), (2000, 128), (500, 129),
                                 (100, 255), (77, 1000), (40, 4099), (5000, 50), (31, 7), (9, 3),
                                 (200, 130), (300, 96), (300, 97)])
def test_autocov_bit_exact(B, n):
    from elfi_b200 import ops
    rs = np.random.RandomState(B + n)
    x = rs.randn(B, n)
    got = ops.autocov(x, lags=(1, 2)).cpu().numpy()
    assert np.array_equal(got[:, 0], np.mean(x[:, 1:] * x[:, :-1], axis=1))
    assert np.array_equal(got[:, 1], np.mean(x[:, 2:] * x[:, :-2], axis=1))
    assert np.array_equal(got[:, 0], o.autocov(x, 1))


@pytest.mark.parametrize('lags', [(1,), (2,), (3,), (4,), (2, 1), (5,), (1, 2, 3, 7)])
def test_autocov_lag_variants(lags):
    from elfi_b200 import ops
    rs = np.random.RandomState(sum(lags))
    x = rs.randn(700, 100)
    got = ops.autocov(x, lags=lags).cpu().numpy()
    for c, lag in enumerate(lags):
        assert np.array_equal(got[:, c], np.mean(x[:, lag:] * x[:, :-lag], axis=1)), lag


@pytest.mark.parametrize('B,n', [(1000, 50), (333, 16), (64, 17), (2000, 128), (500, 129),
                                 (100, 300), (50, 5000), (31, 7), (12, 1),
                                 # contiguous rows with n = 2 mod 4: the 1-D bulk row-group kernel
                                 (1, 50), (33, 2), (37, 18), (4097, 62), (10007, 34), (300001, 50),
                                 (257, 6), (95, 14), (64, 46)])
def test_meanvar_bit_exact(B, n):
    from elfi_b200 import ops
    rs = np.random.RandomState(B * 7 + n)
    y = rs.randn(B, n) * 3 + 1
    got = ops.meanvar(y).cpu().numpy()
    assert np.array_equal(got[:, 0], np.mean(y, axis=1))
    assert np.array_equal(got[:, 1], np.var(y, axis=1))


def test_golden_ma2_pipeline():
    """simulator output -> S1,S2 -> d, all on the device, equals the reference's generate()."""
    from elfi_b200 import ops
    g = load_golden('ma2_generate')
    S = ops.autocov(g['MA2'], lags=(1, 2))
    assert np.array_equal(S[:, 0].cpu().numpy(), g['S1'])
    assert np.array_equal(S[:, 1].cpu().numpy(), g['S2'])
    obs = ops.autocov(g['observed_MA2'], lags=(1, 2))
    assert np.array_equal(obs.cpu().numpy().ravel(), [g['obs_S1'][0], g['obs_S2'][0]])
    d, _ = ops.dist_euclid(S, obs)
    assert np.array_equal(d.cpu().numpy(), g['d'])


def test_golden_gauss_pipeline():
    from elfi_b200 import ops
    g = load_golden('gauss_generate')
    S = ops.meanvar(g['gauss'])
    assert np.array_equal(S[:, 0].cpu().numpy(), g['ss_mean'])
    assert np.array_equal(S[:, 1].cpu().numpy(), g['ss_var'])
    obs = ops.meanvar(g['observed_gauss'])
    d, _ = ops.dist_euclid(S, obs)
    assert np.array_equal(d.cpu().numpy(), g['d'])


def test_full_size_ma2_native():
    """1e6 x 100 MA2-shaped input: sampled rows bit-exact vs the oracle."""
    import torch
    from elfi_b200 import ops
    B, n = 1_000_000, 100
    x = torch.randn(B, n, dtype=torch.float64, device='cuda',
                    generator=torch.Generator(device='cuda').manual_seed(3))
    S = ops.autocov(x, lags=(1, 2))
    rows = torch.cat([torch.arange(0, 1024), torch.arange(B - 1024, B),
                      torch.randint(0, B, (2048,))]).cuda()
    xs = x[rows].cpu().numpy()
    assert np.array_equal(S[rows, 0].cpu().numpy(), o.autocov(xs, 1))
    assert np.array_equal(S[rows, 1].cpu().numpy(), o.autocov(xs, 2))


def test_signed_zero_rows_match_numpy_bits():
    """All-zero / negative-zero rows: the sign of the result follows NumPy's r[k] = a[k] start."""
    from elfi_b200 import ops
    x = np.random.RandomState(5).randn(64, 100)
    x[0] = 0.0
    x[1] = -0.0
    x[2, ::2] = -0.0
    got = ops.autocov(x, lags=(1, 2)).cpu().numpy()
    ref = np.column_stack([np.mean(x[:, 1:] * x[:, :-1], axis=1),
                           np.mean(x[:, 2:] * x[:, :-2], axis=1)])
    assert np.array_equal(got.view(np.int64), ref.view(np.int64))
    y = x[:, :50].copy()
    mv = ops.meanvar(y).cpu().numpy()
    refmv = np.column_stack([np.mean(y, axis=1), np.var(y, axis=1)])
    assert np.array_equal(mv.view(np.int64), refmv.view(np.int64))


@pytest.mark.parametrize('B,n', [(40, 7688), (24, 7696), (16, 8192)])
def test_long_rows_around_the_stack_bound(B, n):
    """7688 terms is the longest run the 6-level pairwise stack of the row-stream kernels holds;
    longer rows must take the thread-per-row kernel (24 levels)."""
    from elfi_b200 import ops
    rs = np.random.RandomState(n)
    x = rs.randn(B, n)
    got = ops.autocov(x, lags=(1, 2)).cpu().numpy()
    assert np.array_equal(got[:, 0], np.mean(x[:, 1:] * x[:, :-1], axis=1))
    assert np.array_equal(got[:, 1], np.mean(x[:, 2:] * x[:, :-2], axis=1))
    mv = ops.meanvar(x).cpu().numpy()
    assert np.array_equal(mv[:, 0], np.mean(x, axis=1))
    assert np.array_equal(mv[:, 1], np.var(x, axis=1))
