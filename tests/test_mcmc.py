"""elfi_b200/mcmc.py against chains produced by the reference's elfi/methods/mcmc.py
(tests/golden/mcmc.npz, generator tests/golden/gen_golden_bolfi.py): same seed -> same chain."""
import numpy as np
import pytest

from conftest import load_golden

PREC = np.linalg.inv(np.array([[1.0, 0.6], [0.6, 0.8]]))
MEAN = np.array([0.5, -0.3])
BOX = (np.array([-1.0, -1.0]), np.array([2.0, 1.0]))


def gauss_logpdf(x):
    d = np.asarray(x) - MEAN
    return -0.5 * d @ PREC @ d


def gauss_grad(x):
    return -PREC @ (np.asarray(x) - MEAN)


def boxed_logpdf(x):
    x = np.asarray(x)
    if np.any(x < BOX[0]) or np.any(x > BOX[1]):
        return -np.inf
    return gauss_logpdf(x)


def boxed_grad(x):
    x = np.asarray(x)
    if np.any(x < BOX[0]) or np.any(x > BOX[1]):
        return np.zeros(2)
    return gauss_grad(x)


def test_nuts_chains_equal_the_reference():
    from elfi_b200 import mcmc
    g = load_golden('mcmc')
    got = mcmc.nuts(300, np.array([1.5, 0.5]), gauss_logpdf, gauss_grad, n_adapt=100, seed=11)
    assert np.array_equal(got, g['nuts_gauss'])
    got = mcmc.nuts(200, np.array([-1.0, 0.2]), gauss_logpdf, gauss_grad, n_adapt=50, max_depth=3,
                    target_prob=0.8, seed=5)
    assert np.array_equal(got, g['nuts_gauss_depth3'])
    got = mcmc.nuts(300, np.array([1.8, 0.9]), boxed_logpdf, boxed_grad, n_adapt=100, seed=3)
    assert np.array_equal(got, g['nuts_boxed'])
    assert got[:, 0].min() >= -1 and got[:, 0].max() <= 2 and np.abs(got[:, 1]).max() <= 1
    got = mcmc.nuts(100, np.array([0.0, 0.0]), gauss_logpdf, gauss_grad, n_adapt=0, stepsize=0.4,
                    seed=7)
    assert np.array_equal(got, g['nuts_fixed_step'])


def test_metropolis_chain_equals_the_reference():
    from elfi_b200 import mcmc
    g = load_golden('mcmc')
    got = mcmc.metropolis(400, np.array([1.5, 0.5]), boxed_logpdf, np.array([0.4, 0.3]), warmup=50,
                          seed=9)
    assert np.array_equal(got, g['metropolis'])


def test_diagnostics_equal_the_reference():
    from elfi_b200 import mcmc
    g = load_golden('mcmc')
    chains = g['chains']
    for k in range(2):
        np.testing.assert_allclose(mcmc.eff_sample_size(chains[:, :, k]), g['ess'][k], rtol=1e-12)
        np.testing.assert_allclose(mcmc.gelman_rubin_statistic(chains[:, :, k]), g['rhat'][k],
                                   rtol=1e-12)
    np.testing.assert_allclose(mcmc.eff_sample_size(chains[0, :, 0]), float(g['ess_single']),
                               rtol=1e-12)


def test_diagnostics_against_stan_values():
    """The known answers of the reference's own test (tests/unit/test_mcmc.py:48-69): chains drawn
    in PyStan, Stan's effective sample size 4.09 and split R-hat 1.714."""
    from elfi_b200 import mcmc
    g = load_golden('mcmc')
    assert np.isclose(mcmc.eff_sample_size(g['stan_chains']), float(g['stan_ess']), atol=0.01)
    assert np.isclose(mcmc.gelman_rubin_statistic(g['stan_chains']), float(g['stan_rhat']),
                      atol=0.01)


def test_bad_initial_point_raises():
    from elfi_b200 import mcmc
    with pytest.raises(ValueError):
        mcmc.nuts(10, np.array([5.0, 5.0]), boxed_logpdf, boxed_grad)
    with pytest.raises(ValueError):
        mcmc.metropolis(10, np.array([5.0, 5.0]), boxed_logpdf, np.array([0.1, 0.1]))


def test_nuts_recovers_the_target_moments():
    from elfi_b200 import mcmc
    chain = mcmc.nuts(3000, np.array([0.0, 0.0]), gauss_logpdf, gauss_grad, n_adapt=500, seed=1)
    kept = chain[500:]
    assert np.all(np.abs(kept.mean(axis=0) - MEAN) < 0.15)
    assert np.all(np.abs(np.cov(kept.T) - np.array([[1.0, 0.6], [0.6, 0.8]])) < 0.25)


def test_lockstep_driver_equals_sequential_chains():
    """run_lockstep answers the pending requests of all chains with one batched evaluation per
    round; every chain still visits exactly the points it visits when run alone."""
    from elfi_b200 import mcmc
    calls = []

    def evaluate(X, with_grad):
        calls.append((len(X), with_grad))
        return (np.array([boxed_logpdf(x) for x in X]),
                np.array([boxed_grad(x) for x in X]) if with_grad else None)

    starts = [np.array([1.5, 0.5]), np.array([0.0, 0.0]), np.array([-0.5, 0.8]), np.array([1.0, -0.5])]
    alone = [mcmc.nuts(120, s0, boxed_logpdf, boxed_grad, n_adapt=40, seed=30 + i)
             for i, s0 in enumerate(starts)]
    together = mcmc.run_lockstep([mcmc.nuts_chain(120, s0, n_adapt=40, seed=30 + i)
                                  for i, s0 in enumerate(starts)], evaluate)
    for a, b in zip(alone, together):
        assert np.array_equal(a, b)
    n_points = sum(k for k, _ in calls)
    assert max(k for k, _ in calls) == 4 and len(calls) < 0.5 * n_points    # rounds are shared
    # Metropolis never asks for gradients
    calls.clear()
    alone = [mcmc.metropolis(200, s0, boxed_logpdf, np.array([0.3, 0.2]), warmup=20, seed=i)
             for i, s0 in enumerate(starts)]
    together = mcmc.run_lockstep([mcmc.metropolis_chain(200, s0, np.array([0.3, 0.2]), warmup=20,
                                                        seed=i) for i, s0 in enumerate(starts)],
                                 evaluate)
    for a, b in zip(alone, together):
        assert np.array_equal(a, b)
    assert not any(g for _, g in calls) and len(calls) == 221
    # a chain that fails raises through the driver
    with pytest.raises(ValueError):
        mcmc.run_lockstep([mcmc.nuts_chain(10, np.array([9.0, 9.0]))], evaluate)
