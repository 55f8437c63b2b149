"""GPU parity: Rejection / SMC / AdaptiveDistanceSMC through the elfi_b200 node + sampler API vs
golden results produced by the reference with the same seeds (tests/golden/gen_golden.py)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _ma2():
    from elfi_b200.examples import ma2
    return ma2.get_model(seed_obs=4)


def test_generate_matches_reference():
    g = load_golden('ma2_generate')
    m = _ma2()
    out = m.generate(1000, ['t1', 't2', 'MA2', 'S1', 'S2', 'd'], seed=123)
    for k in ['t1', 't2', 'MA2']:
        assert np.array_equal(out[k], g[k])
    for k in ['S1', 'S2', 'd']:
        assert np.array_equal(out[k].cpu().numpy(), g[k]), k
    assert np.array_equal(np.asarray(m['S1'].observed.cpu()), g['obs_S1'])


def test_rejection_quantile_config1():
    """BASELINE config #1: MA2 Rejection, batch_size=1000, n_samples=100, quantile 0.01."""
    import elfi_b200 as elfi
    g = load_golden('ma2_rejection_quantile')
    m = _ma2()
    res = elfi.Rejection(m['d'], batch_size=1000, seed=123).sample(100, quantile=0.01, bar=False)
    assert res.n_sim == int(g['n_sim']) == 10000
    assert res.threshold == float(g['threshold']) == 0.10757588712310992
    assert np.array_equal(res.discrepancies, g['out_d'])
    assert np.array_equal(res.samples['t1'], g['out_t1'])
    assert np.array_equal(res.samples['t2'], g['out_t2'])
    assert res.accept_rate == 0.01
    assert len(np.unique(res.discrepancies)) == 100


def test_rejection_threshold_mode():
    import elfi_b200 as elfi
    g = load_golden('ma2_rejection_threshold')
    m = _ma2()
    res = elfi.Rejection(m['d'], batch_size=1000, seed=123).sample(150, threshold=0.2, bar=False)
    assert res.n_sim == int(g['n_sim'])
    assert res.n_batches == int(g['n_batches'])
    assert res.threshold == float(g['threshold'])
    assert np.array_equal(res.discrepancies, g['out_d'])
    assert np.array_equal(res.samples['t1'], g['out_t1'])
    assert np.all(res.discrepancies <= 0.2)


def test_rejection_nsim_mode_with_extra_outputs():
    import elfi_b200 as elfi
    g = load_golden('ma2_rejection_nsim')
    m = _ma2()
    res = elfi.Rejection(m['d'], batch_size=500, seed=7, output_names=['S1', 'S2']).sample(
        64, n_sim=3000, bar=False)
    assert res.n_sim == int(g['n_sim'])
    for k in ['d', 't1', 't2', 'S1', 'S2']:
        assert np.array_equal(res.outputs[k], g['out_' + k]), k


def _check_smc(res, g, exact_first=True):
    assert res.n_sim == int(g['n_sim'])
    assert len(res.populations) == int(g['n_pops'])
    for i, pop in enumerate(res.populations):
        pre = 'pop{}_'.format(i)
        assert pop.n_sim == int(g[pre + 'n_sim']), i
        for k in pop.outputs:
            ref = g[pre + 'out_' + k]
            if i == 0 and exact_first:
                assert np.array_equal(pop.outputs[k], ref), (i, k)
            else:
                np.testing.assert_allclose(pop.outputs[k], ref, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(pop.weights, g[pre + 'weights'], rtol=1e-5)   # north_star tol
        np.testing.assert_allclose(pop.cov, g[pre + 'cov'], rtol=1e-6)
        np.testing.assert_allclose(pop.threshold, float(g[pre + 'threshold']), rtol=1e-7)
    np.testing.assert_allclose(res.weights, g['weights'], rtol=1e-5)
    np.testing.assert_allclose(res.threshold, float(g['threshold']), rtol=1e-7)


def test_smc_quantiles_ma2():
    import elfi_b200 as elfi
    g = load_golden('ma2_smc_quantiles')
    res = elfi.SMC(_ma2()['d'], batch_size=1000, seed=123).sample(200, quantiles=[.5, .5, .5],
                                                                  bar=False)
    _check_smc(res, g)
    assert res.n_sim == 4000
    np.testing.assert_allclose(res.weights[:3], [0.28398978528812313, 0.27910050379622714,
                                                 0.2537994934197125], rtol=1e-5)


def test_smc_thresholds_ma2():
    import elfi_b200 as elfi
    g = load_golden('ma2_smc_thresholds')
    res = elfi.SMC(_ma2()['d'], batch_size=1000, seed=20).sample(150, thresholds=[.6, .3, .15],
                                                                 bar=False)
    _check_smc(res, g)


def test_smc_quantiles_gauss():
    import elfi_b200 as elfi
    from elfi_b200.examples import gauss
    g = load_golden('gauss_smc_quantiles')
    m = gauss.get_model(n_obs=50, seed_obs=3)
    res = elfi.SMC(m['d'], batch_size=1000, seed=9).sample(300, quantiles=[.3, .3, .3], bar=False)
    _check_smc(res, g)


def test_adaptive_distance_smc_ma2():
    import elfi_b200 as elfi
    g = load_golden('ma2_adaptive_distance_smc')
    m = _ma2()
    m['d'].become(elfi.AdaptiveDistance(m['S1'], m['S2']))
    res = elfi.AdaptiveDistanceSMC(m['d'], batch_size=500, seed=11).sample(
        100, rounds=3, quantile=0.5, bar=False)
    assert res.n_sim == int(g['n_sim'])
    for i, pop in enumerate(res.populations):
        pre = 'pop{}_'.format(i)
        assert pop.n_sim == int(g[pre + 'n_sim'])
        np.testing.assert_allclose(pop.adaptive_distance_w, g[pre + 'w'], rtol=1e-9)
        for k in ['d', 't1', 't2', 'S1', 'S2']:
            np.testing.assert_allclose(pop.outputs[k], g[pre + 'out_' + k], rtol=1e-6, atol=1e-9,
                                       err_msg='pop {} output {}'.format(i, k))
        np.testing.assert_allclose(pop.weights, g[pre + 'weights'], rtol=1e-5)
        np.testing.assert_allclose(pop.threshold, float(g[pre + 'threshold']), rtol=1e-7)


def test_seed_determinism_and_difference():
    """tests/functional/test_consistency.py of the reference: same seed -> same result."""
    import elfi_b200 as elfi
    m = _ma2()
    a = elfi.Rejection(m['d'], batch_size=500, seed=1).sample(50, quantile=0.05, bar=False)
    b = elfi.Rejection(m['d'], batch_size=500, seed=1).sample(50, quantile=0.05, bar=False)
    c = elfi.Rejection(m['d'], batch_size=500, seed=2).sample(50, quantile=0.05, bar=False)
    assert np.array_equal(a.samples_array, b.samples_array)
    assert not np.array_equal(a.samples_array, c.samples_array)


def test_statistical_posterior_means():
    """tests/functional/test_inference.py:16-55 of the reference: means near (0.6, 0.2)."""
    import elfi_b200 as elfi
    m = _ma2()
    res = elfi.Rejection(m['d'], batch_size=20000, seed=3).sample(1000, quantile=0.01, bar=False)
    assert abs(res.sample_means['t1'] - 0.6) < 0.05
    assert abs(res.sample_means['t2'] - 0.2) < 0.05
    assert res.n_sim == 100000


def test_adaptive_distance_smc_gnk_config5_model():
    """BASELINE config #5's model at a size the reference can run: g-and-k, (B, n_obs) order
    statistics on the device, AdaptiveDistance, AdaptiveDistanceSMC."""
    import elfi_b200 as elfi
    from elfi_b200.examples import gnk
    g = load_golden('gnk_adaptive_distance_smc')
    m = gnk.get_adaptive_model(n_obs=64, seed=7)
    res = elfi.AdaptiveDistanceSMC(m['d'], batch_size=400, seed=13).sample(100, rounds=2,
                                                                          quantile=0.5, bar=False)
    assert res.n_sim == int(g['n_sim'])
    assert len(res.populations) == int(g['n_pops'])
    for i, pop in enumerate(res.populations):
        pre = 'pop{}_'.format(i)
        assert pop.n_sim == int(g[pre + 'n_sim'])
        np.testing.assert_allclose(pop.adaptive_distance_w, g[pre + 'w'], rtol=1e-9)
        for k in ('d', 'A', 'B', 'g', 'k', 'ss_sorted'):
            np.testing.assert_allclose(pop.outputs[k], g[pre + 'out_' + k], rtol=1e-6, atol=1e-9,
                                       err_msg='pop {} output {}'.format(i, k))
        np.testing.assert_allclose(pop.weights, g[pre + 'weights'], rtol=1e-5)
