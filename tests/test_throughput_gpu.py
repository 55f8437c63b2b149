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
