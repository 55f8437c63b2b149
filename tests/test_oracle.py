"""Pin the CPU oracle: (i) against the SciPy/NumPy routines the reference calls, (ii) against
golden vectors produced by the reference itself (tests/golden/gen_golden.py)."""
import numpy as np
import pytest
from scipy.spatial.distance import cdist

import elfi_oracle as o
from conftest import load_golden


@pytest.mark.parametrize('B,D', [(1000, 128), (777, 2), (100, 1), (513, 255), (64, 17), (5, 256)])
def test_cdist_bit_exact(B, D):
    rs = np.random.RandomState(B + D)
    S = rs.randn(B, D)
    obs = rs.randn(1, D)
    w = rs.rand(D) + 0.1
    assert np.array_equal(o.cdist_euclid(S, obs), cdist(S, obs, 'euclidean').ravel())
    assert np.array_equal(o.cdist_euclid(S, obs, w), cdist(S, obs, 'euclidean', w=w).ravel())
    assert np.array_equal(o.cdist_euclid(S, obs, threads=3), cdist(S, obs).ravel())


def test_nested_distance_matches_cdist_columns():
    rs = np.random.RandomState(3)
    S = rs.randn(300, 40)
    obs = rs.randn(1, 40)
    ws = [None, rs.rand(40) + .5, rs.rand(40) + .5]
    got = o.nested_distance(S, obs, ws)
    for k, w in enumerate(ws):
        ref = cdist(S, obs, 'euclidean', w=None if w is None else w ** 2).ravel()
        assert np.array_equal(got[:, k], ref)


@pytest.mark.parametrize('n', [1, 5, 7, 8, 9, 50, 99, 100, 127, 128, 129, 255, 256, 1000, 1023])
def test_numpy_pairwise_order(n):
    rs = np.random.RandomState(n)
    x = rs.randn(200, n)
    m, v = o.meanvar(x)
    assert np.array_equal(m, np.mean(x, axis=1))
    assert np.array_equal(v, np.var(x, axis=1))
    for lag in (1, 2):
        if n > lag:
            assert np.array_equal(o.autocov(x, lag), np.mean(x[:, lag:] * x[:, :-lag], axis=1))


def test_numpy_reduce_starts_from_identity():
    """np.add.reduce = 0.0 + pairwise_sum: an all negative-zero run sums to +0.0."""
    x = np.random.RandomState(2).randn(8, 100)
    x[0] = -0.0
    x[1, ::2] = -0.0
    x[2] = 0.0
    m, v = o.meanvar(x)
    assert np.array_equal(m.view(np.int64), np.mean(x, axis=1).view(np.int64))
    assert np.array_equal(v.view(np.int64), np.var(x, axis=1).view(np.int64))
    for lag in (1, 2):
        ref = np.mean(x[:, lag:] * x[:, :-lag], axis=1)
        assert np.array_equal(o.autocov(x, lag).view(np.int64), ref.view(np.int64))


def test_golden_ma2_forward():
    g = load_golden('ma2_generate')
    x = g['MA2']
    assert np.array_equal(o.autocov(x, 1), g['S1'])
    assert np.array_equal(o.autocov(x, 2), g['S2'])
    obs = np.array([g['obs_S1'][0], g['obs_S2'][0]])
    assert np.array_equal(o.cdist_euclid(np.column_stack([g['S1'], g['S2']]), obs), g['d'])
    assert np.array_equal(o.autocov(g['observed_MA2'], 1), g['obs_S1'])


def test_golden_gauss_forward():
    g = load_golden('gauss_generate')
    m, v = o.meanvar(g['gauss'])
    assert np.array_equal(m, g['ss_mean'])
    assert np.array_equal(v, g['ss_var'])
    obs = np.array([g['obs_ss_mean'][0], g['obs_ss_var'][0]])
    assert np.array_equal(o.cdist_euclid(np.column_stack([m, v]), obs), g['d'])


def test_golden_weighted_quantile_and_var():
    g = load_golden('weighted_quantile')
    for a, qw, qu in zip(g['alphas'], g['q_w'], g['q_unw']):
        assert o.weighted_sample_quantile(g['x'], a, g['w']) == qw
        assert o.weighted_sample_quantile(g['x'], a) == qu
    g = load_golden('weighted_var')
    assert np.array_equal(o.weighted_var(g['x'], g['w']), g['var_w'])
    assert np.array_equal(o.weighted_var(g['x']), g['var_unw'])


def test_golden_gm_logpdf():
    g = load_golden('gm_logpdf')
    got = o.gm_logpdf(g['x'], g['means'], g['cov'], g['weights'])
    np.testing.assert_allclose(got, g['logpdf'], rtol=1e-10, atol=1e-12)
    g = load_golden('gm_logpdf_fullcov')
    got = o.gm_logpdf(g['x'], g['means'], g['cov'], g['weights'])
    np.testing.assert_allclose(got, g['logpdf'], rtol=1e-10, atol=1e-12)


def test_merge_batch_matches_topn():
    rs = np.random.RandomState(0)
    n, B = 50, 400
    samples = {'d': np.ones(n + B) * np.inf, 't': np.empty(n + B)}
    alld, allt = [], []
    for _ in range(4):
        batch = {'d': rs.rand(B), 't': rs.randn(B)}
        o.merge_batch(samples, batch, n, None, 'd')
        alld.append(batch['d'])
        allt.append(batch['t'])
    alld = np.concatenate(alld)
    allt = np.concatenate(allt)
    order = np.argsort(alld)[:n]
    assert np.array_equal(samples['d'][:n], alld[order])
    assert np.array_equal(samples['t'][:n], allt[order])


def test_welford_matches_std():
    rs = np.random.RandomState(1)
    store = [0, 0, 0]
    data = []
    for _ in range(3):
        chunk = rs.randn(200, 5) * [1, 2, 3, 4, 5]
        data.append(chunk)
        store, scale = o.welford_add(store, chunk)
    np.testing.assert_allclose(scale, np.std(np.vstack(data), axis=0), rtol=1e-12)


def test_gp_oracle_vs_sklearn():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C
    rs = np.random.RandomState(0)
    X = rs.uniform(-2, 2, (150, 2))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] ** 2 + 0.05 * rs.randn(150)
    s2, ell, b, noise = 1.3, 0.7, 0.4, 0.01
    L, alpha = o.gp_fit(X, y, s2, ell, b, noise, jitter=0.0)
    xs = rs.uniform(-2, 2, (64, 2))
    mu, var = o.gp_predict(xs, X, L, alpha, s2, ell, b)
    gp = GaussianProcessRegressor(kernel=C(s2) * RBF(ell) + C(b), alpha=noise, optimizer=None)
    gp.fit(X, y)
    m2, sd2 = gp.predict(xs, return_std=True)
    np.testing.assert_allclose(mu.ravel(), m2, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(var.ravel(), sd2 ** 2, rtol=1e-6, atol=1e-9)


def test_gp_gradients_finite_difference():
    rs = np.random.RandomState(2)
    X = rs.uniform(-1, 1, (60, 2))
    y = np.cos(2 * X[:, 0]) * X[:, 1]
    s2, ell, b, noise = 0.8, 0.5, 0.2, 0.02
    L, alpha = o.gp_fit(X, y, s2, ell, b, noise)
    x0 = rs.uniform(-1, 1, (5, 2))
    gm, gv = o.gp_predictive_gradients(x0, X, L, alpha, s2, ell, b)
    h = 1e-6
    for j in range(2):
        e = np.zeros(2)
        e[j] = h
        mp, vp = o.gp_predict(x0 + e, X, L, alpha, s2, ell, b)
        mm, vm = o.gp_predict(x0 - e, X, L, alpha, s2, ell, b)
        np.testing.assert_allclose(gm[:, j], ((mp - mm) / (2 * h)).ravel(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(gv[:, j], ((vp - vm) / (2 * h)).ravel(), rtol=1e-4, atol=1e-7)


def test_golden_kliep():
    """oracle KLIEP restatement vs the reference's DensityRatioEstimation (fixed sigma)."""
    g = load_golden('kliep')
    alpha, max_ratio = o.kliep_fit(g['x'], g['y'], g['wx'], g['wy'], sigma=float(g['sigma']))
    np.testing.assert_allclose(max_ratio, float(g['max_ratio']), rtol=1e-12)
    theta = g['x'][:100]
    d2 = ((g['x'][:, None, :] - theta[None, :, :]) ** 2).sum(-1)
    ratios = np.exp(-0.5 * d2 / float(g['sigma']) ** 2) @ alpha
    np.testing.assert_allclose(ratios, g['ratios'], rtol=1e-11)


@pytest.mark.parametrize('metric,kw', [('sqeuclidean', {}), ('cityblock', {}), ('chebyshev', {}),
                                       ('minkowski', {'p': 3.0}), ('minkowski', {'p': 0.5}),
                                       ('minkowski', {'p': 1.5})])
def test_other_cdist_metrics_are_sequential(metric, kw):
    """SciPy accumulates these metrics left to right like 'euclidean': the sequential C
    restatement is bit-identical (Minkowski: same libm pow), which is what the device kernels
    of elfi_b200_dist_metric_thr_f64 reproduce."""
    from scipy.spatial.distance import cdist
    rs = np.random.RandomState(3)
    for B, D in [(500, 128), (64, 3), (200, 17), (1, 1)]:
        S, obs = rs.randn(B, D), rs.randn(1, D)
        ref = cdist(S, obs, metric, **kw).ravel()
        got = o.cdist_metric(S, obs, metric, kw.get('p', 2.0))
        if metric == 'minkowski':
            np.testing.assert_allclose(got, ref, rtol=2e-16 * 8)
        else:
            assert np.array_equal(got, ref), (metric, B, D)


def test_seuclidean_two_sum_order():
    """SciPy's 'seuclidean' loop is NOT one left-to-right sum: even and odd columns accumulate
    separately, meet, then the odd-D tail is added.  The C restatement (and the device kernel of
    elfi_b200_dist_seuclidean_thr_f64 that follows it) is bit-identical to the installed SciPy;
    a sequential sum and the w = 1/V route are not."""
    from scipy.spatial.distance import cdist
    rs = np.random.RandomState(5)
    sequential_differs = False
    for D in list(range(1, 41)) + [63, 64, 65, 127, 128, 129, 255, 1000]:
        B = 200
        S = rs.randn(B, D) * rs.uniform(0.1, 100)
        obs = rs.randn(1, D)
        V = rs.uniform(0.01, 50, D)
        ref = cdist(S, obs, 'seuclidean', V=V).ravel()
        assert np.array_equal(o.cdist_seuclidean(S, obs, V), ref), D
        if D >= 16 and not np.array_equal(o.cdist_euclid(S, obs, w=1.0 / V), ref):
            sequential_differs = True
    assert sequential_differs
    # several observed rows use the same loop per pair
    S, O, V = rs.randn(50, 9), rs.randn(3, 9), rs.uniform(0.1, 3, 9)
    ref = cdist(S, O, 'seuclidean', V=V)
    for k in range(3):
        assert np.array_equal(o.cdist_seuclidean(S, O[k:k + 1], V), ref[:, k])
