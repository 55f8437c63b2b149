"""Throughput mode (device-side priors / simulator / proposals, Philox streams): statistical
parity with the reference's host path and exact self-consistency checks."""
import numpy as np
import pytest
import scipy.stats as ss

import elfi_oracle as o

pytestmark = pytest.mark.gpu


def test_fused_summaries_equal_materialised_data():
    """sim_ma2 with fused autocov == autocov(materialised X) bit for bit, for several n_obs."""
    from elfi_b200 import ops
    rs = np.random.RandomState(0)
    for n_obs in (100, 3, 8, 9, 10, 17, 130, 257, 1000):
        B = 777
        t1, t2 = rs.uniform(-1, 1, B), rs.uniform(-0.5, 0.5, B)
        X, S = ops.sim_ma2(t1, t2, n_obs, seed=42, offset=5, want_data=True, want_summaries=True)
        X = X.cpu().numpy()
        assert np.array_equal(S[:, 0].cpu().numpy(), o.autocov(X, 1)), n_obs
        assert np.array_equal(S[:, 1].cpu().numpy(), o.autocov(X, 2)), n_obs
        _, S2 = ops.sim_ma2(t1, t2, n_obs, seed=42, offset=5)
        assert np.array_equal(S2.cpu().numpy(), S.cpu().numpy())
        # rows are a pure function of (seed, offset + row): sharding-invariant
        Xs, _ = ops.sim_ma2(t1[100:], t2[100:], n_obs, seed=42, offset=105, want_data=True,
                            want_summaries=False)
        assert np.array_equal(Xs.cpu().numpy(), X[100:])


def test_simulator_statistics():
    """MA2 second moments: var = 1 + t1^2 + t2^2, lag-1 cov = t1 + t1 t2, lag-2 cov = t2."""
    from elfi_b200 import ops
    B = 200000
    t1, t2 = np.full(B, 0.6), np.full(B, 0.2)
    X, S = ops.sim_ma2(t1, t2, 100, seed=7, want_data=True)
    X = X.cpu().numpy()
    assert abs(X.mean()) < 0.01
    assert abs(X.var() - (1 + 0.36 + 0.04)) < 0.01
    S = S.cpu().numpy()
    assert abs(S[:, 0].mean() - (0.6 + 0.12) * 99 / 99) < 0.01
    assert abs(S[:, 1].mean() - 0.2) < 0.01
    # normality of the innovations: x / sd ~ N(0, 1)
    z = X[:2000, 50] / np.sqrt(1.4)
    assert ss.kstest(z, 'norm').pvalue > 1e-3


def test_device_prior_matches_host_distribution():
    from elfi_b200 import ops
    from elfi_b200.examples import ma2
    t1, t2 = ops.prior_ma2(100000, seed=3)
    t1, t2 = t1.cpu().numpy(), t2.cpu().numpy()
    rs = np.random.RandomState(0)
    h1 = ma2.CustomPrior1.rvs(2, size=100000, random_state=rs)
    h2 = ma2.CustomPrior2.rvs(h1, 1, size=100000, random_state=rs)
    assert ss.ks_2samp(t1, h1).pvalue > 1e-3
    assert ss.ks_2samp(t2, h2).pvalue > 1e-3
    assert np.all(np.abs(t1) <= 2) and np.all(t2 >= -1 + np.abs(t1) - 1e-12) and np.all(t2 <= 1)
    only1 = ops.prior_ma2(1000, seed=3, which='t1').cpu().numpy()
    assert np.array_equal(only1, t1[:1000])
    cond = ops.prior_ma2(0, seed=3, t1=only1, which='t2').cpu().numpy()
    assert np.array_equal(cond, t2[:1000])


def test_logprior_matches_model_prior():
    from elfi_b200 import ops
    from elfi_b200.examples import ma2
    from elfi_b200.samplers import ModelPrior
    rs = np.random.RandomState(1)
    theta = np.column_stack([rs.uniform(-2.5, 2.5, 2000), rs.uniform(-1.5, 1.5, 2000)])
    with np.errstate(divide='ignore'):
        ref = ModelPrior(ma2.get_model(seed_obs=1)).logpdf(theta)
    got = ops.logprior_ma2(theta).cpu().numpy()
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    fin = np.isfinite(ref)
    np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-12, atol=1e-12)


def test_gm_rvs_moments_and_support():
    from elfi_b200 import ops
    rs = np.random.RandomState(2)
    means = np.column_stack([rs.uniform(0.3, 0.9, 500), rs.uniform(0.0, 0.4, 500)])
    w = rs.rand(500)
    cov = np.array([[0.02, 0.005], [0.005, 0.01]])
    x = ops.gm_rvs(means, cov, w, 200000, seed=9, support=1).cpu().numpy()
    wn = w / w.sum()
    mix_mean = wn @ means
    np.testing.assert_allclose(x.mean(0), mix_mean, atol=5e-3)
    mix_cov = cov + (means - mix_mean).T @ ((means - mix_mean) * wn[:, None])
    np.testing.assert_allclose(np.cov(x.T), mix_cov, atol=2e-3)
    assert np.all(np.abs(x[:, 0]) < 2) and np.all(x[:, 1] >= -1 + np.abs(x[:, 0])) and np.all(x[:, 1] <= 1)
    # support rejection: a mixture straddling the boundary never leaves the support
    x2 = ops.gm_rvs(np.array([[1.9, 0.95]]), np.eye(2) * 0.05, None, 50000, seed=1, support=1)
    x2 = x2.cpu().numpy()
    assert np.all(np.abs(x2[:, 0]) < 2) and np.all(x2[:, 1] >= -1 + np.abs(x2[:, 0])) and np.all(x2[:, 1] <= 1)


def test_rejection_throughput_mode_statistics():
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_device_model(seed_obs=4)
    res = elfi.Rejection(m['d'], batch_size=100000, seed=3).sample(2000, quantile=0.005, bar=False)
    assert res.n_sim == 400000
    assert abs(res.sample_means['t1'] - 0.6) < 0.05 and abs(res.sample_means['t2'] - 0.2) < 0.05
    again = elfi.Rejection(m['d'], batch_size=100000, seed=3).sample(2000, quantile=0.005, bar=False)
    assert np.array_equal(res.samples_array, again.samples_array)      # deterministic given the seed


def test_smc_throughput_mode_statistics():
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_device_model(seed_obs=4)
    smc = elfi.SMC(m['d'], batch_size=50000, seed=5, device_proposal=ma2.DeviceProposal)
    res = smc.sample(5000, quantiles=[0.2, 0.3, 0.3], bar=False)
    assert len(res.populations) == 3
    thr = [p.threshold for p in res.populations]
    assert thr[0] > thr[1] > thr[2]
    means = res.sample_means_array
    assert abs(means[0] - 0.6) < 0.05 and abs(means[1] - 0.2) < 0.05, means
    # reference host path on the same task agrees statistically
    mh = ma2.get_model(seed_obs=4)
    ref = elfi.SMC(mh['d'], batch_size=50000, seed=5).sample(5000, quantiles=[0.2, 0.3, 0.3],
                                                             bar=False)
    assert abs(ref.sample_means_array[0] - means[0]) < 0.03
    assert abs(ref.sample_means_array[1] - means[1]) < 0.03


def test_gauss_fused_summaries_and_prior():
    from elfi_b200 import ops
    rs = np.random.RandomState(0)
    for n_obs in (50, 1, 7, 8, 9, 129, 500):
        B = 555
        mu, sigma = rs.uniform(-1, 9, B), rs.uniform(0.1, 3, B)
        Y, S = ops.sim_gauss(mu, sigma, n_obs, seed=11, offset=3, want_data=True)
        Y = Y.cpu().numpy()
        assert np.array_equal(S[:, 0].cpu().numpy(), np.mean(Y, axis=1)), n_obs
        assert np.array_equal(S[:, 1].cpu().numpy(), np.var(Y, axis=1)), n_obs
    prm = [-1.0, 10.0, 0.01, 10.0]
    mu, sigma = ops.prior_gauss(200000, seed=5, prm=prm)
    mu, sigma = mu.cpu().numpy(), sigma.cpu().numpy()
    assert ss.kstest(mu, ss.uniform(-1, 10).cdf).pvalue > 1e-3
    assert ss.kstest(sigma, ss.truncnorm(0.01, 10).cdf).pvalue > 1e-3
    theta = np.column_stack([rs.uniform(-2, 10, 3000), rs.uniform(-1, 11, 3000)])
    with np.errstate(divide='ignore'):
        ref = ss.uniform.logpdf(theta[:, 0], -1, 10) + ss.truncnorm.logpdf(theta[:, 1], 0.01, 10)
    got = ops.logprior_gauss(theta, prm).cpu().numpy()
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    np.testing.assert_allclose(got[np.isfinite(ref)], ref[np.isfinite(ref)], rtol=1e-11)
    x = ops.gm_rvs(np.array([[8.9, 0.02]]), np.eye(2) * 0.3, None, 20000, seed=2, support=2,
                   box=([-1.0, 0.01], [9.0, 10.0])).cpu().numpy()
    assert x[:, 0].min() >= -1 and x[:, 0].max() <= 9 and x[:, 1].min() >= 0.01


def test_gauss_smc_throughput_mode_statistics():
    """config #3's model (Gaussian noise, SMC-ABC) on the device vs the host-RNG path."""
    import elfi_b200 as elfi
    from elfi_b200.examples import gauss
    m, proposal = gauss.get_device_model(n_obs=50, seed_obs=3)
    res = elfi.SMC(m['d'], batch_size=50000, seed=4, device_proposal=proposal).sample(
        4000, quantiles=[0.2, 0.3, 0.3], bar=False)
    mh = gauss.get_model(n_obs=50, seed_obs=3)
    ref = elfi.SMC(mh['d'], batch_size=50000, seed=4).sample(4000, quantiles=[0.2, 0.3, 0.3],
                                                            bar=False)
    a, b = res.sample_means_array, ref.sample_means_array
    assert abs(a[0] - b[0]) < 0.05 and abs(a[1] - b[1]) < 0.05, (a, b)
    assert abs(a[0] - 4.0) < 0.3 and abs(a[1] - 0.4) < 0.3


def _gnk_q(prm, z, c=0.8):
    A, B, g, k = prm
    return A + B * (1 + c * ((1 - np.exp(-g * z)) / (1 + np.exp(-g * z)))) * (1 + z**2)**k * z


def test_gnk_simulator_distribution_and_streams():
    """sim_gnk draws follow the g-and-k quantile function (gnk.py:60-66): the fraction of draws
    below Q(Phi^-1(p)) is p; rows are a pure function of (seed, offset + row, column)."""
    from elfi_b200 import ops
    prm = (3.0, 1.0, 2.0, 0.5)
    B, n_obs = 4000, 50
    cols = [np.full(B, v) for v in prm]
    Y = ops.sim_gnk(*cols, n_obs=n_obs, seed=21).cpu().numpy()
    assert Y.shape == (B, n_obs) and np.all(np.isfinite(Y))
    N = Y.size
    for p in np.arange(1, 8) / 8.0:
        frac = np.mean(Y <= _gnk_q(prm, ss.norm.ppf(p)))
        assert abs(frac - p) < 5 * np.sqrt(p * (1 - p) / N) + 1e-9, (p, frac)
    # every row has its own stream: row means differ, columns are uncorrelated
    assert np.unique(Y[:, 0]).size == B
    assert abs(np.corrcoef(Y[:, 0], Y[:, 1])[0, 1]) < 0.08
    # sharding invariance and independence of n_obs (odd n_obs: last column written alone)
    Y2 = ops.sim_gnk(*[c[40:100] for c in cols], n_obs=n_obs, seed=21, offset=40).cpu().numpy()
    assert np.array_equal(Y2, Y[40:100])
    Y3 = ops.sim_gnk(*[c[:64] for c in cols], n_obs=51, seed=21).cpu().numpy()
    assert np.array_equal(Y3[:, :50], Y[:64])
    Y4 = ops.sim_gnk(*[c[:64] for c in cols], n_obs=1, seed=21).cpu().numpy()
    assert np.array_equal(Y4[:, 0], Y[:64, 0])
    # per-row parameters are honoured: location shift by A
    A2 = cols[0].copy()
    A2[::2] += 100.0
    Y5 = ops.sim_gnk(A2, *cols[1:], n_obs=n_obs, seed=21).cpu().numpy()
    np.testing.assert_allclose(Y5[::2] - 100.0, Y[::2], rtol=0, atol=1e-10)
    assert np.array_equal(Y5[1::2], Y[1::2])


def test_logprior_box_matches_scipy():
    from elfi_b200 import ops
    rs = np.random.RandomState(4)
    theta = rs.uniform(-1, 11, (5000, 4))
    theta[0] = [0.0, 10.0, 5.0, 5.0]            # boundary points are inside (scipy: closed support)
    theta[1] = [np.nan, 1.0, 1.0, 1.0]
    with np.errstate(divide='ignore'):
        ref = ss.uniform.logpdf(theta, 0, 10).sum(axis=1)
    ref[1] = -np.inf                            # NaN is outside for the acceptance test
    got = ops.logprior_box(theta, 0.0, 10.0).cpu().numpy()
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    np.testing.assert_allclose(got[np.isfinite(ref)], ref[np.isfinite(ref)], rtol=1e-14)
    got2 = ops.logprior_box(theta[:, :2], [0.0, -1.0], [10.0, 2.0]).cpu().numpy()
    with np.errstate(divide='ignore'):
        ref2 = ss.uniform.logpdf(theta[:, 0], 0, 10) + ss.uniform.logpdf(theta[:, 1], -1, 2)
    ref2[1] = -np.inf
    assert np.array_equal(np.isfinite(got2), np.isfinite(ref2))
    np.testing.assert_allclose(got2[np.isfinite(ref2)], ref2[np.isfinite(ref2)], rtol=1e-14)


def test_gnk_adaptive_distance_smc_throughput_mode_statistics():
    """config #5's model with priors, simulator and proposals on the device vs the host-RNG path
    (statistical parity: the two use different random streams)."""
    import elfi_b200 as elfi
    from elfi_b200.examples import gnk
    m, proposal = gnk.get_device_model(n_obs=64, seed=7)
    res = elfi.AdaptiveDistanceSMC(m['d'], batch_size=4000, seed=13,
                                   device_proposal=proposal).sample(500, rounds=3, quantile=0.5,
                                                                    bar=False)
    mh = gnk.get_adaptive_model(n_obs=64, seed=7)
    ref = elfi.AdaptiveDistanceSMC(mh['d'], batch_size=4000, seed=13).sample(500, rounds=3,
                                                                             quantile=0.5, bar=False)
    assert res.n_samples == 500 and len(res.populations) == len(ref.populations) == 3
    thr = [pop.threshold for pop in res.populations]
    assert all(np.isfinite(thr))
    for pop in res.populations:
        assert np.all(np.isfinite(pop.weights)) and pop.weights.sum() > 0
        for name in ('A', 'B', 'g', 'k'):
            v = pop.outputs[name]
            assert v.min() >= 0.0 and v.max() <= 10.0
    a, b = res.sample_means_array, ref.sample_means_array
    for i, tol in enumerate((1.0, 1.0, 1.5, 1.5)):
        assert abs(a[i] - b[i]) < tol, (a, b)
