"""Gaussian noise model (mirror of elfi/examples/gauss.py, 1-d case) with device summaries."""
from functools import partial

import numpy as np
import scipy.stats as ss

from .. import model as em
from .. import ops


def gauss(mu, sigma, n_obs=50, batch_size=1, random_state=None):
    """elfi/examples/gauss.py:11-35."""
    batches_mu = np.asanyarray(mu).reshape((-1, 1))
    batches_sigma = np.asanyarray(sigma).reshape((-1, 1))
    return ss.norm.rvs(loc=batches_mu, scale=batches_sigma, size=(batch_size, n_obs),
                       random_state=random_state)


def ss_mean(y):
    """np.mean(y, axis=1) on the device (elfi/examples/gauss.py:142-156)."""
    return ops.meanvar(y)[:, 0]


def ss_var(y):
    """np.var(y, axis=1) on the device (elfi/examples/gauss.py:159-173)."""
    return ops.meanvar(y)[:, 1]


def get_model(n_obs=50, true_params=None, seed_obs=None):
    """elfi/examples/gauss.py:75-139 (nd_mean=False)."""
    if true_params is None:
        true_params = [4, .4]
    fn_simulator = partial(gauss, n_obs=n_obs)
    y_obs = fn_simulator(*true_params, n_obs=n_obs, random_state=np.random.RandomState(seed_obs))
    m = em.new_model()
    eps_prior = 5
    priors = [em.Prior('uniform', true_params[0] - eps_prior, 2 * eps_prior, model=m, name='mu'),
              em.Prior('truncnorm', np.amax([.01, true_params[1] - eps_prior]), 2 * eps_prior,
                       model=m, name='sigma')]
    em.Simulator(fn_simulator, *priors, observed=y_obs, name='gauss')
    sumstats = [em.Summary(ss_mean, m['gauss'], name='ss_mean'),
                em.Summary(ss_var, m['gauss'], name='ss_var')]
    em.Distance('euclidean', *sumstats, name='d')
    return m
