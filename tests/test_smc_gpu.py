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
