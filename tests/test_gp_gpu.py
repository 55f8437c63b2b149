"""GPU parity: GP surrogate / LCBSC / BOLFI vs the oracle (SciPy Cholesky restatement of
gpy_regression.py:127-160, 206-218, cross-checked against scikit-learn in test_oracle.py)."""
import numpy as np
import pytest

import elfi_oracle as o

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star: GP posterior mean / var within 1e-5 relative


def _data(n, p=2, seed=0):
    rs = np.random.RandomState(seed)
    lo, hi = np.array([-2.0, -1.0, 0.0][:p]), np.array([2.0, 1.0, 3.0][:p])
    X = rs.uniform(lo, hi, (n, p))
    y = np.log(0.05 + np.sum((X - 0.3) ** 2, axis=1)) + 0.1 * rs.randn(n)
    return X, y, list(zip(lo, hi))


def _model(n, p=2, seed=0):
    from elfi_b200.bo import GPyRegression
    X, y, bounds = _data(n, p, seed)
    names = ['t{}'.format(i) for i in range(p)]
    gp = GPyRegression(names, bounds=dict(zip(names, bounds)))
    gp.update(X, y)
    return gp, X, y


@pytest.mark.parametrize('n,p', [(20, 2), (64, 2), (65, 1), (300, 2), (777, 3), (2000, 2)])
def test_gp_predict_matches_oracle(n, p):
    from elfi_b200.bo import JITTER
    gp, X, y = _model(n, p, seed=n)
    h = gp.hyperparameters
    L, alpha = o.gp_fit(X, y, h['kernel_var'], h['lengthscale'], h['bias_var'], h['noise_var'],
                        jitter=JITTER)
    rs = np.random.RandomState(1)
    xq = np.vstack([X[:50], rs.uniform(-2, 2, (1500, p))])
    mu, var = gp.predict(xq, noiseless=True)
    mu_o, var_o = o.gp_predict(xq, X, L, alpha, h['kernel_var'], h['lengthscale'], h['bias_var'])
    np.testing.assert_allclose(mu, mu_o, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(var, var_o, rtol=RTOL, atol=1e-9)
    mu2, var2 = gp.predict(xq[:10])
    np.testing.assert_allclose(var2, var_o[:10] + h['noise_var'], rtol=RTOL, atol=1e-9)
    assert mu.shape == (len(xq), 1) and var.shape == (len(xq), 1)


def test_default_hyperparameters_follow_reference_heuristics():
    """gpy_regression.py:242-284: the kernel starts from GPy's unit values; the heuristics from the
    bounds and the first data parameterise the Gamma priors (Gamma.from_EV(v, v): shape v, rate 1)
    and the noise variance."""
    gp, X, y = _model(50)
    h = gp.hyperparameters
    assert h['kernel_var'] == 1.0 and h['lengthscale'] == 1.0 and h['bias_var'] == 1.0
    assert h['noise_var'] == pytest.approx(np.max(y) ** 2 / 100.)
    kernel_var = (np.max(y) / 3.) ** 2
    assert gp._priors['lengthscale'] == (pytest.approx((2.0 - (-2.0)) / 3.), 1.0)
    assert gp._priors['kernel_var'] == (pytest.approx(kernel_var), 1.0)
    assert gp._priors['bias_var'] == (pytest.approx(kernel_var / 4.), 1.0)


def test_gp_gradients_match_oracle():
    from elfi_b200.bo import JITTER
    gp, X, y = _model(400, 2, seed=5)
    h = gp.hyperparameters
    L, alpha = o.gp_fit(X, y, h['kernel_var'], h['lengthscale'], h['bias_var'], h['noise_var'],
                        jitter=JITTER)
    xq = np.random.RandomState(2).uniform(-1, 1, (12, 2))
    gm, gv = gp.predictive_gradients(xq)
    gm_o, gv_o = o.gp_predictive_gradients(xq, X, L, alpha, h['kernel_var'], h['lengthscale'],
                                           h['bias_var'])
    np.testing.assert_allclose(gm, gm_o, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(gv, gv_o, rtol=RTOL, atol=1e-9)


def test_chunked_triangular_products_are_consistent():
    """gp_trimv_kernel handles the query points in chunks of up to 16 right-hand sides (templates
    for 4, 8, 16): one call with 80 points, calls of 20 and calls of 16 agree, and the whitened
    vectors reproduce the predictive variance of the GEMM path."""
    from elfi_b200 import device as dev
    for n in (77, 400, 1300):
        gp, X, y = _model(n, 2, seed=n)
        xq = np.random.RandomState(4).uniform(-1, 1, (80, 2))
        gm_a, gv_a = gp.predictive_gradients(xq)
        parts = [gp.predictive_gradients(xq[i:i + 20]) for i in range(0, 80, 20)]
        gm_b = np.vstack([q[0] for q in parts])
        gv_b = np.vstack([q[1] for q in parts])
        np.testing.assert_allclose(gm_b, gm_a, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(gv_b, gv_a, rtol=1e-10, atol=1e-12)
        _, T_a = gp.whiten(xq)
        T_b = np.vstack([dev.to_host(gp.whiten(xq[i:i + 16])[1]) for i in range(0, 80, 16)])
        np.testing.assert_allclose(T_b, dev.to_host(T_a), rtol=1e-10, atol=1e-13)
        mu, var = gp.predict(xq[:5], noiseless=True)
        h = gp.hyperparameters
        np.testing.assert_allclose(var.ravel(), h['kernel_var'] + h['bias_var'] -
                                   np.sum(T_b[:5] ** 2, axis=1), rtol=1e-9, atol=1e-12)


def test_lcbsc_value_and_gradient():
    from elfi_b200.bo import LCBSC, JITTER
    gp, X, y = _model(256, 2, seed=8)
    h = gp.hyperparameters
    acq = LCBSC(gp, exploration_rate=10, seed=0)
    t = 10
    beta = o.lcbsc_beta(t, 2, 10)
    assert acq._beta(t) == pytest.approx(beta)
    L, alpha = o.gp_fit(X, y, h['kernel_var'], h['lengthscale'], h['bias_var'], h['noise_var'],
                        jitter=JITTER)
    xq = np.random.RandomState(3).uniform(-1, 1, (9, 2))
    mu, var = o.gp_predict(xq, X, L, alpha, h['kernel_var'], h['lengthscale'], h['bias_var'])
    np.testing.assert_allclose(acq.evaluate(xq, t), o.lcbsc(mu, var, beta), rtol=RTOL, atol=1e-9)
    gm, gv = o.gp_predictive_gradients(xq, X, L, alpha, h['kernel_var'], h['lengthscale'],
                                       h['bias_var'])
    np.testing.assert_allclose(acq.evaluate_gradient(xq, t), o.lcbsc_gradient(var, gm, gv, beta),
                               rtol=RTOL, atol=1e-8)


def test_config4_grid_properties():
    """BASELINE config #4: 2000 evidence, 1e5-point grid.  Oracle on a 4000-point sample of the
    grid; whole-grid properties (finite, 0 <= var <= prior var, LCBSC = mean - sqrt(beta var))."""
    from elfi_b200.bo import LCBSC, JITTER
    gp, X, y = _model(2000, 2, seed=0)
    h = gp.hyperparameters
    g1, g2 = np.meshgrid(np.linspace(-2, 2, 400), np.linspace(-1, 1, 250))
    grid = np.column_stack([g1.ravel(), g2.ravel()])
    acq = LCBSC(gp, exploration_rate=10, seed=0)
    beta = acq._beta(10)
    mean, var, a = gp.predict_device(grid, noiseless=True, beta=beta)
    mean, var, a = mean.cpu().numpy(), var.cpu().numpy(), a.cpu().numpy()
    assert np.all(np.isfinite(mean)) and np.all(np.isfinite(var))
    assert var.min() > 0 and var.max() <= h['kernel_var'] + h['bias_var'] + 1e-9
    np.testing.assert_allclose(a, mean - np.sqrt(beta * var), rtol=1e-12)
    L, alpha = o.gp_fit(X, y, h['kernel_var'], h['lengthscale'], h['bias_var'], h['noise_var'],
                        jitter=JITTER)
    idx = np.random.RandomState(0).choice(len(grid), 4000, replace=False)
    mu_o, var_o = o.gp_predict(grid[idx], X, L, alpha, h['kernel_var'], h['lengthscale'],
                               h['bias_var'])
    np.testing.assert_allclose(mean[idx], mu_o.ravel(), rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(var[idx], var_o.ravel(), rtol=RTOL, atol=1e-9)


def test_bad_pivot_raises():
    from elfi_b200.bo import GPyRegression
    gp = GPyRegression(['a'], bounds={'a': (0, 1)}, noise_var=1e-30)
    X = np.zeros((40, 1))                     # identical points -> singular Ky without noise
    with pytest.raises(np.linalg.LinAlgError):
        gp._X = X
        gp._Y = np.ones((40, 1))
        gp._hyper = dict(kernel_var=1.0, lengthscale=1.0, bias_var=0.0, noise_var=-1e-3)
        gp._fit()


def test_optimize_improves_marginal_likelihood():
    gp, X, y = _model(150, 2, seed=3)
    before = gp.log_posterior_hyper()
    gp.optimize()
    assert gp.log_posterior_hyper() >= before - 1e-9


def test_bolfi_ma2_smoke():
    """tests/functional/test_inference.py:136-190 of the reference, shortened: BOLFI on MA2."""
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_model(seed_obs=4)
    log_d = elfi.Operation(np.log, m['d'], name='log_d')
    bolfi = elfi.BOLFI(log_d, batch_size=5, initial_evidence=20, update_interval=10,
                       bounds={'t1': (-2, 2), 't2': (-1, 1)}, acq_noise_var=0.1, seed=1)
    post = bolfi.fit(n_evidence=60, bar=False)
    assert bolfi.target_model.n_evidence == 60
    res = bolfi.extract_result()
    assert abs(res.x_min['t1'][0] - 0.6) < 0.6 and abs(res.x_min['t2'][0] - 0.2) < 0.6
    lp = post.logpdf(np.array([[0.6, 0.2], [-1.5, 0.9]]))
    assert lp[0] > lp[1]


def test_bolfi_ma2_reference_bounds():
    """The reference's own BOLFI test (tests/functional/test_inference.py:136-190) at its own
    size and error bound: 300 evidence points, |x_min - true| < 0.2 for both parameters,
    continuation keeps the acquired points, the maximum-likelihood point of the extracted
    posterior and the NUTS sample mean are within 0.2 as well."""
    import elfi_b200 as elfi
    from elfi_b200.bo import minimize
    from elfi_b200.examples import ma2
    m = ma2.get_model(n_obs=100, true_params=[.6, .2], seed_obs=4)
    log_d = elfi.Operation(np.log, m['d'], name='log_d')
    bolfi = elfi.BOLFI(log_d, initial_evidence=20, update_interval=10, batch_size=5,
                       bounds={'t1': (-2, 2), 't2': (-1, 1)}, acq_noise_var=.1, seed=1)
    n = 300
    res = bolfi.infer(n, bar=False)
    assert bolfi.target_model.n_evidence == n
    acq_x = bolfi.target_model.X.copy()
    assert abs(res.x_min['t1'][0] - 0.6) < 0.2 and abs(res.x_min['t2'][0] - 0.2) < 0.2, res.x_min
    res = bolfi.infer(n + 10, bar=False)
    assert bolfi.target_model.n_evidence == n + 10
    assert np.array_equal(bolfi.target_model.X[:n, :], acq_x)
    post = bolfi.extract_posterior()
    post_ml = minimize(lambda x: -post._unnormalized_loglikelihood(x), post.model.bounds,
                       grad=lambda x: -post._gradient_unnormalized_loglikelihood(x),
                       prior=post.prior, n_start_points=post.n_inits, maxiter=post.max_opt_iters,
                       random_state=np.random.RandomState(0))[0]
    assert abs(post_ml[0] - 0.6) < 0.2 and abs(post_ml[1] - 0.2) < 0.2, post_ml
    n_samples, n_chains = 400, 4
    sample = bolfi.sample(n_samples, n_chains=n_chains)
    assert len(sample.samples['t1']) == n_samples // 2 * n_chains
    assert abs(np.mean(sample.samples['t1']) - 0.6) < 0.2
    assert abs(np.mean(sample.samples['t2']) - 0.2) < 0.2
