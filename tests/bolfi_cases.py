"""Bodies of the BOLFI posterior / acquisition / sampling tests, shared by the CPU-double and
GPU collections.  Goldens: tests/golden/bolfi_posterior.npz (reference formulas on a duck GP
backed by the oracle; generator tests/golden/gen_golden_bolfi.py)."""
import numpy as np

from conftest import load_golden

RTOL = 1e-5   # north_star tolerance for GP posterior quantities


def _fixture_gp():
    from elfi_b200.bo import GPyRegression
    g = load_golden('bolfi_posterior')
    names = ['t1', 't2']
    gp = GPyRegression(names, bounds={'t1': (-2, 2), 't2': (-1, 1)})
    gp.update(g['X'], g['y'][:, None])
    gp._hyper = dict(zip(('kernel_var', 'lengthscale', 'bias_var', 'noise_var'),
                         (float(v) for v in g['hyper'])))
    gp._fit()
    return g, gp


def _prior():
    from elfi_b200.examples import ma2
    from elfi_b200.samplers import ModelPrior
    return ModelPrior(ma2.get_model(seed_obs=4))


def case_prior_gradient():
    g = load_golden('bolfi_posterior')
    prior = _prior()
    with np.errstate(all='ignore'):
        got = prior.gradient_logpdf(g['pts'])
        np.testing.assert_allclose(prior.logpdf(g['pts']), g['prior_logpdf'], rtol=1e-12)
    np.testing.assert_allclose(got, g['prior_grad'], rtol=1e-9, atol=1e-12)
    assert prior.gradient_logpdf(g['pts'][1]).shape == (2,)


def case_posterior_matches_reference():
    from elfi_b200.bo import BolfiPosterior
    g, gp = _fixture_gp()
    post = BolfiPosterior(gp, threshold=float(g['threshold']), prior=_prior())
    with np.errstate(all='ignore'):
        loglik = post._unnormalized_loglikelihood(g['pts'])
        gradlik = post._gradient_unnormalized_loglikelihood(g['pts'])
        logpdf = post.logpdf(g['pts'])
        grad = post.gradient_logpdf(g['pts'])
    fin = np.isfinite(g['loglik'])
    assert np.array_equal(np.isfinite(loglik), fin)            # -inf outside the GP bounds
    np.testing.assert_allclose(loglik[fin], g['loglik'][fin], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(gradlik, g['gradlik'], rtol=1e-4, atol=1e-7)
    fin = np.isfinite(g['logpdf'])
    assert np.array_equal(np.isfinite(logpdf), fin)            # and outside the prior support
    np.testing.assert_allclose(logpdf[fin], g['logpdf'][fin], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(grad, g['grad'], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(post.logpdf(g['pts'][1]), float(g['single_logpdf']), rtol=RTOL)
    np.testing.assert_allclose(post.gradient_logpdf(g['pts'][1]), g['single_grad'], rtol=1e-4,
                               atol=1e-7)
    assert np.ndim(post.logpdf(g['pts'][1])) == 0 and post.gradient_logpdf(g['pts'][1]).shape == (2,)
    # default threshold: minimum of the GP mean found by the multi-start minimiser
    post_default = BolfiPosterior(gp, prior=_prior(), seed=0)
    np.testing.assert_allclose(post_default.threshold, float(g['default_threshold']), rtol=1e-4)


def case_maxvar_matches_reference():
    from elfi_b200.bo import MaxVar
    g, gp = _fixture_gp()
    acq = MaxVar(model=gp, prior=_prior(), quantile_eps=0.05, noise_var=0.1, seed=1)
    acq._update_eps()
    np.testing.assert_allclose(acq.eps, float(g['maxvar_eps']), rtol=1e-12)
    got = acq.evaluate(g['maxvar_pts'])
    assert got.shape == g['maxvar'].shape
    np.testing.assert_allclose(got, g['maxvar'], rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(acq.evaluate_gradient(g['maxvar_pts']), g['maxvar_grad'], rtol=1e-3,
                               atol=1e-9)
    x = acq.acquire(3)
    assert x.shape == (3, 2) and np.all(x[0] == x[1])
    assert -2 <= x[0, 0] <= 2 and -1 <= x[0, 1] <= 1
    assert acq.evaluate(x[:1])[0, 0] >= got.max() * 0.999      # a maximiser beats the probe points


def case_lcbsc_acquire_matches_reference():
    """LCBSC.evaluate / evaluate_gradient / acquire / _add_noise and bo.utils.minimize against the
    reference's own classes run on a duck GP (tests/golden/lcbsc_acquire.npz; acquisition.py:129-301,
    bo/utils.py:40-111): same start points, same optimiser, same RandomState consumption, so the
    acquired points agree to the accuracy of the GP posterior itself."""
    from elfi_b200.bo import LCBSC, GPyRegression, minimize, minimize_lockstep
    g = load_golden('lcbsc_acquire')
    gp = GPyRegression(['t1', 't2'], bounds={'t1': (-2, 2), 't2': (-1, 1)})
    gp.update(g['X'], g['y'][:, None])
    gp._hyper = dict(zip(('kernel_var', 'lengthscale', 'bias_var', 'noise_var'),
                         (float(v) for v in g['hyper'])))
    gp._fit()
    prior = _prior()
    lc = LCBSC(model=gp, prior=prior, noise_var=0.1, exploration_rate=10, seed=1, n_inits=10,
               max_opt_iters=1000)
    for t in (0, 4, 40):
        np.testing.assert_allclose(lc._beta(t), float(g['beta_t{}'.format(t)]), rtol=1e-14)
        val = lc.evaluate(g['pts'], t)
        assert val.shape == g['value_t{}'.format(t)].shape
        np.testing.assert_allclose(val, g['value_t{}'.format(t)], rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(lc.evaluate_gradient(g['pts'], t), g['grad_t{}'.format(t)],
                                   rtol=1e-4, atol=1e-7)
    # acquire: multi-start L-BFGS-B from prior draws, n copies, truncated-normal noise; two calls in
    # a row continue the same RandomState exactly as the reference does
    for lockstep in (True, False):
        lc = LCBSC(model=gp, prior=prior, noise_var=0.1, exploration_rate=10, seed=1, n_inits=10,
                   max_opt_iters=1000)
        lc.lockstep = lockstep
        np.testing.assert_allclose(lc.acquire(3, t=4), g['acquired_t4'], rtol=2e-4, atol=2e-4)
        np.testing.assert_allclose(lc.acquire(2, t=5), g['acquired_t5'], rtol=2e-4, atol=2e-4)
    quiet = LCBSC(model=gp, prior=prior, noise_var=0, exploration_rate=10, seed=3, n_inits=6)
    xq = quiet.acquire(2, t=2)
    np.testing.assert_allclose(xq, g['acquired_quiet'], rtol=2e-4, atol=2e-4)
    assert np.array_equal(xq[0], xq[1])                     # zero noise: n copies of the minimiser
    by_name = LCBSC(model=gp, prior=None, noise_var={'t1': 0.2, 't2': 0.0}, delta=0.2, seed=5)
    assert by_name.exploration_rate == 5.0
    xd = by_name.acquire(4, t=1)
    np.testing.assert_allclose(xd, g['acquired_dict'], rtol=2e-4, atol=2e-4)
    assert np.all(xd[:, 1] == xd[0, 1]) and len(np.unique(xd[:, 0])) == 4   # noise on t1 only
    noisy = LCBSC(model=gp, prior=prior, noise_var=0.5, seed=9)
    out = noisy._add_noise(g['noise_in'].copy())
    assert np.array_equal(out, g['noise_out'])              # pure host RNG: bit-identical
    assert np.all((out[:, 0] >= -2) & (out[:, 0] <= 2) & (out[:, 1] >= -1) & (out[:, 1] <= 1))

    # bo.utils.minimize on an analytic function: prior and uniform start points
    def quad(x):
        x = np.atleast_2d(x)
        return (x[:, 0] - 0.3) ** 2 + 3 * (x[:, 1] + 0.4) ** 2 + 0.5 * np.sin(5 * x[:, 0])

    def quad_grad(x):
        x = np.atleast_2d(x)
        return np.column_stack([2 * (x[:, 0] - 0.3) + 2.5 * np.cos(5 * x[:, 0]), 6 * (x[:, 1] + 0.4)])
    bounds = [(-2, 2), (-1, 1)]
    loc, val = minimize(quad, bounds, grad=quad_grad, prior=prior, n_start_points=7,
                        random_state=np.random.RandomState(4))
    assert np.array_equal(loc, g['min_loc']) and val == float(g['min_val'])
    loc, val = minimize(quad, bounds, grad=quad_grad, prior=None, n_start_points=5,
                        random_state=np.random.RandomState(6))
    assert np.array_equal(loc, g['min_loc_uniform']) and val == float(g['min_val_uniform'])
    loc, val = minimize_lockstep(lambda X: (quad(X), quad_grad(X)), bounds, prior=prior,
                                 n_start_points=7, random_state=np.random.RandomState(4))
    assert np.array_equal(loc, g['min_loc']) and val == float(g['min_val'])


def case_hyper_objective_matches_sklearn():
    """The objective of GPyRegression.optimize(): its log marginal likelihood (device Cholesky)
    and the gradient in log-parameters that L-BFGS-B sees, against scikit-learn's
    GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True) for the same
    kernel sigma^2 RBF(l) + b + noise (theta = log parameters); GPy itself (SCG + paramz
    transforms, gpy_regression.py:317-323) is not installable here."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, WhiteKernel
    from elfi_b200.bo import JITTER, GPyRegression
    import scipy.stats as ss
    rs = np.random.RandomState(12)
    n = 90
    X = np.column_stack([rs.uniform(-2, 2, n), rs.uniform(-1, 1, n)])
    y = np.log(0.05 + (X[:, 0] - 0.6) ** 2 + 2 * (X[:, 1] - 0.2) ** 2) + 0.2 * rs.randn(n)
    gp = GPyRegression(['t1', 't2'], bounds={'t1': (-2, 2), 't2': (-1, 1)})
    gp.update(X, y[:, None])
    names = ['kernel_var', 'lengthscale', 'bias_var', 'noise_var']
    for hyper in ([1.0, 1.0, 1.0, 0.2], [2.3, 0.6, 0.5, 0.07], [0.4, 1.7, 3.0, 0.5]):
        h = dict(zip(names, hyper))
        kernel = ConstantKernel(h['kernel_var']) * RBF(h['lengthscale']) + \
            ConstantKernel(h['bias_var']) + WhiteKernel(h['noise_var'] + JITTER)
        sk = GaussianProcessRegressor(kernel=kernel, alpha=0.0, optimizer=None).fit(X, y)
        # sklearn orders theta as log[kernel_var, lengthscale, bias_var, noise]
        want, want_grad = sk.log_marginal_likelihood(sk.kernel_.theta, eval_gradient=True)
        got = gp.log_marginal_likelihood(h)
        np.testing.assert_allclose(got, want, rtol=1e-9)

        def lml(logh):
            return gp.log_marginal_likelihood(dict(zip(names, np.exp(logh))))
        x0, eps = np.log(hyper), 1e-5
        fd = np.array([(lml(x0 + eps * e) - lml(x0 - eps * e)) / (2 * eps) for e in np.eye(4)])
        # d/dlog(noise + jitter) vs d/dlog(noise): identical up to jitter / noise ~ 1e-7
        np.testing.assert_allclose(fd, want_grad, rtol=2e-5, atol=1e-6)
        # on top: the reference's Gamma log-priors (gpy_regression.py:267-280) and, as in GPy, the
        # log Jacobian log(1 - e^-x) of the Logexp transform of every priored parameter
        prior = sum(ss.gamma.logpdf(h[k], a=a, scale=1.0 / b) + np.log(-np.expm1(-h[k]))
                    for k, (a, b) in gp._priors.items())
        np.testing.assert_allclose(gp.log_posterior_hyper(h), got + prior, rtol=1e-12)
    gp._fit()


def case_expintvar_matches_reference():
    """ExpIntVar on a grid: the candidate-dependent loss equals the reference's (which factorises
    Ky per evaluation) at probe points; the acquired point is at least as good as the reference's."""
    from elfi_b200.bo import ExpIntVar, GPyRegression
    g = load_golden('expintvar')
    gp = GPyRegression(['t1', 't2'], bounds={'t1': (-2, 2), 't2': (-1, 1)})
    gp.update(g['X'], g['y'][:, None])
    gp._hyper = dict(zip(('kernel_var', 'lengthscale', 'bias_var', 'noise_var'),
                         (float(v) for v in g['hyper'])))
    gp._fit()
    acq = ExpIntVar(model=gp, prior=_prior(), quantile_eps=0.05, integration='grid', d_grid=0.4,
                    noise_var=0.1, seed=1, n_inits=5, max_opt_iters=100)
    assert np.array_equal(acq.points_int, g['grid'])
    acq._prepare(0)
    np.testing.assert_allclose(acq.eps, float(g['eps']), rtol=1e-12)
    np.testing.assert_allclose(acq.omegas_int, g['omegas'], rtol=1e-12)
    np.testing.assert_allclose(acq.phi_int, g['phi_int'], rtol=1e-5, atol=1e-12)
    loss = acq.evaluate(g['pts'])
    np.testing.assert_allclose(loss, g['loss'], rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(acq.evaluate(g['pts'][1]), g['single'], rtol=1e-4, atol=1e-12)
    x = acq.acquire(2, t=0)
    assert x.shape == (2, 2) and np.all(x[0] == x[1])
    assert acq.evaluate(x[:1])[0] <= float(g['loss'][-1]) * (1 + 1e-3) + 1e-12
    # the covariance pieces themselves: symmetric, and the self-covariance is the variance
    pts = g['pts'][:4]
    wh = gp.whiten(pts)
    cov = gp.cross_covariance(wh, wh).cpu().numpy()
    np.testing.assert_allclose(cov, cov.T, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.diag(cov), gp.predict(pts, noiseless=True)[1].ravel(),
                               rtol=1e-7, atol=1e-10)
    # importance-sampling variant runs end to end
    imp = ExpIntVar(model=gp, prior=_prior(), quantile_eps=0.05, integration='importance',
                    n_samples_imp=20, n_samples=60, sampler='metropolis', noise_var=0.1, seed=2,
                    n_inits=2, max_opt_iters=20)
    xi = imp.acquire(1, t=0)
    assert xi.shape == (1, 2) and np.isfinite(xi).all()


def case_other_acquisitions():
    from elfi_b200.bo import RandMaxVar, UniformAcquisition
    g, gp = _fixture_gp()
    u = UniformAcquisition(model=gp, seed=3).acquire(500)
    assert u.shape == (500, 2) and u[:, 0].min() >= -2 and u[:, 0].max() <= 2
    assert np.abs(u[:, 1]).max() <= 1 and abs(u[:, 0].mean()) < 0.3
    for sampler in ('nuts', 'metropolis'):
        acq = RandMaxVar(model=gp, prior=_prior(), quantile_eps=0.05, sampler=sampler,
                         n_samples=30, seed=2)
        x = acq.acquire(1)
        assert x.shape == (1, 2) and np.isfinite(x).all() and acq.evaluate(x)[0, 0] > 0
        xs = acq.acquire(4)
        assert xs.shape == (4, 2)


def case_bolfi_sample():
    """BOLFI.fit + sample on MA2 (tests/functional/test_inference.py:136-190 of the reference,
    shortened): chains stay inside the bounds and concentrate near the true parameters."""
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_model(seed_obs=4)
    log_d = elfi.Operation(np.log, m['d'], name='log_d')
    bolfi = elfi.BOLFI(log_d, batch_size=5, initial_evidence=30, update_interval=10,
                       bounds={'t1': (-2, 2), 't2': (-1, 1)}, acq_noise_var=0.1, seed=1)
    bolfi.fit(n_evidence=80, bar=False)
    res = bolfi.sample(200, n_chains=2, info_freq=1000)
    assert res.chains.shape == (2, 200, 2) and res.n_chains == 2 and res.warmup == 100
    assert res.samples_array.shape == (200, 2) and res.n_sim == 80
    assert np.all(np.abs(res.samples['t1']) <= 2) and np.all(np.abs(res.samples['t2']) <= 1)
    means = res.sample_means_array
    assert abs(means[0] - 0.6) < 0.6 and abs(means[1] - 0.2) < 0.6, means
    again = bolfi.sample(200, n_chains=2, info_freq=1000)
    assert np.array_equal(again.chains, res.chains)            # seeded by get_sub_seed(seed, chain)
    met = bolfi.sample(300, n_chains=2, algorithm='metropolis', sigma_proposals={'t1': 0.3, 't2': 0.2})
    assert met.chains.shape == (2, 300, 2) and met.samples_array.shape == (300, 2)
    init = np.array([[0.5, 0.1], [0.7, 0.3], [0.4, 0.2]])
    custom = bolfi.sample(60, n_chains=3, initials=init, threshold=float(res.threshold))
    assert custom.chains.shape == (3, 60, 2)
    for bad in (dict(algorithm='gibbs'), dict(initials=init[:2], n_chains=3)):
        try:
            bolfi.sample(10, **bad)
        except ValueError:
            continue
        raise AssertionError('expected ValueError for {}'.format(bad))


def case_incremental_factor_update():
    """GPyRegression(incremental=True): appending evidence by rank-b updates gives the factor,
    alpha and predictions of a refit; a refit happens when the padded size or the
    hyper-parameters change."""
    from elfi_b200.bo import GPyRegression
    rs = np.random.RandomState(5)
    bounds = {'t1': (-2, 2), 't2': (-1, 1)}
    X = np.column_stack([rs.uniform(-2, 2, 200), rs.uniform(-1, 1, 200)])
    y = np.log(0.05 + (X[:, 0] - 0.6) ** 2 + 2 * (X[:, 1] - 0.2) ** 2) + 0.1 * rs.randn(200)
    inc = GPyRegression(['t1', 't2'], bounds=bounds, incremental=True)
    ref = GPyRegression(['t1', 't2'], bounds=bounds)
    inc.update(X[:100], y[:100, None])
    ref.update(X[:100], y[:100, None])
    grid = np.column_stack([rs.uniform(-2, 2, 50), rs.uniform(-1, 1, 50)])
    step = [1, 5, 3, 16, 2, 1]              # crosses the 128-row padding boundary on the way
    n = 100
    for b in step:
        inc.update(X[n:n + b], y[n:n + b, None])
        ref.update(X[n:n + b], y[n:n + b, None])
        n += b
        assert inc.n_evidence == ref.n_evidence == n and inc._factor['n'] == n
        np.testing.assert_allclose(inc._factor['alpha'].cpu().numpy(),
                                   ref._factor['alpha'].cpu().numpy(), rtol=1e-6, atol=1e-8)
        for a, r in zip(inc.predict(grid), ref.predict(grid)):
            np.testing.assert_allclose(a, r, rtol=1e-7, atol=1e-9)
        for a, r in zip(inc.predictive_gradients(grid[:5]), ref.predictive_gradients(grid[:5])):
            np.testing.assert_allclose(a, r, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(inc.log_marginal_likelihood(), ref.log_marginal_likelihood(),
                                   rtol=1e-9)
    Wi = inc._factor['W'].cpu().numpy()[:n, :n]
    Wr = ref._factor['W'].cpu().numpy()[:n, :n]
    np.testing.assert_allclose(Wi, Wr, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(inc._factor['U'].cpu().numpy()[:n, :n], Wi.T, rtol=0, atol=0)
    inc.update(X[n:n + 4], y[n:n + 4, None], optimize=True)      # new hyper-parameters: refit
    ref.update(X[n:n + 4], y[n:n + 4, None], optimize=True)
    for a, r in zip(inc.predict(grid), ref.predict(grid)):
        np.testing.assert_allclose(a, r, rtol=1e-7, atol=1e-9)
