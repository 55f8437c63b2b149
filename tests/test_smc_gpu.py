"""GPU parity: SMC population arithmetic vs the oracle and the reference goldens."""
import numpy as np
import pytest

import elfi_oracle as o
from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_weighted_var_golden():
    from elfi_b200 import ops
    g = load_golden('weighted_var')
    np.testing.assert_allclose(ops.weighted_var(g['x'], g['w']), g['var_w'], rtol=1e-12)
    np.testing.assert_allclose(ops.weighted_var(g['x']), g['var_unw'], rtol=1e-12)
    V1, V2, xbar, s2 = ops.weighted_stats(g['x'], g['w'])
    np.testing.assert_allclose(V1, g['w'].sum(), rtol=1e-13)
    np.testing.assert_allclose(xbar, np.average(g['x'], weights=g['w'], axis=0), rtol=1e-12)


def test_gm_logpdf_golden():
    from elfi_b200 import ops
    for name in ('gm_logpdf', 'gm_logpdf_fullcov'):
        g = load_golden(name)
        got = ops.gm_logpdf(g['x'], g['means'], g['cov'], g['weights']).cpu().numpy()
        # 1e-5 relative on the density (north_star tolerance on SMC weights); we are far inside
        np.testing.assert_allclose(np.exp(got), np.exp(g['logpdf']), rtol=1e-7)
        np.testing.assert_allclose(got, g['logpdf'], rtol=0, atol=1e-7)


@pytest.mark.parametrize('p', [1, 2, 3, 4, 6])
def test_gm_logpdf_vs_oracle(p):
    from elfi_b200 import ops
    rs = np.random.RandomState(p)
    M, N = 3000, 1777
    means = rs.randn(M, p)
    w = rs.rand(M)
    A = rs.randn(p, p)
    cov = A @ A.T / p + np.eye(p) * 0.1
    x = rs.randn(N, p) * 1.5
    x[:5] += 40.0                                   # far tail: density underflows like the reference
    ref = o.gm_logpdf(x, means, cov, w)
    got = ops.gm_logpdf(x, means, cov, w).cpu().numpy()
    finite = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), finite)
    np.testing.assert_allclose(got[finite], ref[finite], rtol=0, atol=1e-7)


@pytest.mark.parametrize('p', [1, 2, 3, 4])
def test_gm_logpdf_mixed_vs_oracle(p):
    """elfi_b200_gm_logpdf_mixed_f64 (2^f from the fp32 special-function unit; throughput mode):
    densities within 1e-6 of the oracle -- 10x inside the 1e-5 tolerance on SMC weights -- and the
    same underflow pattern as the fp64 path."""
    from elfi_b200 import ops
    rs = np.random.RandomState(10 + p)
    M, N = 3000, 1777
    means = rs.randn(M, p)
    w = rs.rand(M)
    w[::11] = 0.0
    A = rs.randn(p, p)
    cov = A @ A.T / p + np.eye(p) * 0.1
    x = rs.randn(N, p) * 1.5
    x[:5] += 40.0
    ref = o.gm_logpdf(x, means, cov, w)
    got = ops.gm_logpdf(x, means, cov, w, mixed=True).cpu().numpy()
    exact = ops.gm_logpdf(x, means, cov, w).cpu().numpy()
    finite = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), finite)
    np.testing.assert_allclose(np.exp(got[finite] - ref[finite]), 1.0, rtol=1e-6)
    np.testing.assert_allclose(np.exp(exact[finite] - ref[finite]), 1.0, rtol=1e-7)


def test_fast_exp_accuracy_through_weights():
    """w = exp(logprior - logq) within 1e-5 relative of the oracle (north_star tolerance)."""
    from elfi_b200 import ops
    rs = np.random.RandomState(0)
    means = rs.randn(5000, 2) * 0.3
    wprev = rs.rand(5000)
    cov = np.diag([0.02, 0.01])
    x = means[rs.randint(0, 5000, 2000)] + rs.multivariate_normal([0, 0], cov, 2000)
    logprior = rs.randn(2000) * 0.1
    ref = np.exp(logprior - o.gm_logpdf(x, means, cov, wprev))
    got = ops.smc_weights(logprior, ops.gm_logpdf(x, means, cov, wprev)).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-6)


@pytest.mark.parametrize('B,D', [(1000, 2), (4096, 256), (100003, 33), (64, 1)])
def test_colmoments(B, D):
    from elfi_b200 import ops
    rs = np.random.RandomState(B + D)
    S = rs.randn(B, D) * rs.uniform(0.1, 50, D) + rs.uniform(-100, 100, D)
    mean, m2 = ops.colmoments(S)
    np.testing.assert_allclose(mean, S.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(np.sqrt(m2 / B), S.std(axis=0), rtol=1e-10)


@pytest.mark.parametrize('B,D,K', [(4096, 256, 1), (100003, 33, 3), (5000, 128, 5), (31, 16, 2),
                                   (70000, 256, 8), (1000, 2, 2), (2048, 40, 17)])
def test_fused_distance_and_colmoments(B, D, K):
    """elfi_b200_dist_euclid_mom_f64: nested distances + acceptance bit-identical to the plain
    kernel, column moments of the same read equal to NumPy's (and to the stand-alone pass)."""
    import elfi_oracle as o
    from elfi_b200 import ops
    rs = np.random.RandomState(B + D + K)
    S = rs.randn(B, D) * rs.uniform(0.1, 50, D) + rs.uniform(-100, 100, D)
    obs = rs.randn(1, D) * 20
    W = rs.uniform(0.01, 2.0, (K, D))
    W[0] = 1.0
    ref = np.column_stack([o.cdist_euclid(S, obs, w=W[k]) for k in range(K)])
    thr = np.quantile(ref, 0.6, axis=0)
    d, idx, mom = ops.dist_euclid(S, obs, w=W, thresholds=thr, moments=True)
    d2, idx2 = ops.dist_euclid(S, obs, w=W, thresholds=thr)
    assert np.array_equal(d.cpu().numpy(), ref)
    assert np.array_equal(d.cpu().numpy(), d2.cpu().numpy())
    assert np.array_equal(idx.cpu().numpy(), idx2.cpu().numpy())
    assert np.array_equal(idx.cpu().numpy(), np.nonzero(np.all(ref <= thr, axis=1))[0])
    mean, m2 = mom.cpu().numpy()
    np.testing.assert_allclose(mean, S.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(np.sqrt(m2 / B), S.std(axis=0), rtol=1e-10)
    mean_s, m2_s = ops.colmoments(S)
    np.testing.assert_allclose(mean, mean_s, rtol=1e-12)
    np.testing.assert_allclose(m2, m2_s, rtol=1e-9)
    # a row-strided view (leading dimension > D) and device thresholds
    from elfi_b200 import device as dev
    big = dev.to_device(rs.randn(3000, 96))
    view = big[:, :64]
    obs2 = rs.randn(64)
    Wv = rs.uniform(0.5, 1.5, (2, 64))
    thr_dev = dev.to_device(np.array([12.0, 12.5]))
    dv, (iv, nv), mv = ops.dist_euclid(view, obs2, w=Wv, thresholds=thr_dev, sync=False, moments=True)
    host = view.cpu().numpy()
    refv = np.column_stack([o.cdist_euclid(np.ascontiguousarray(host), obs2, w=Wv[k]) for k in range(2)])
    assert np.array_equal(dv.cpu().numpy(), refv)
    k = int(nv.item())
    assert np.array_equal(iv[:k].cpu().numpy(), np.nonzero(np.all(refv <= [12.0, 12.5], axis=1))[0])
    np.testing.assert_allclose(mv.cpu().numpy()[0], host.mean(axis=0), rtol=1e-11, atol=1e-13)
