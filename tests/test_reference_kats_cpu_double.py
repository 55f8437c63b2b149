"""Known-answer checks restated from the reference's own unit tests (tests/unit/test_utils.py),
run against this implementation on the CPU test double: the same inputs and the same expected
values, through the device operators' wrappers."""
import numpy as np
import pytest
import scipy.stats as ss

pytestmark = pytest.mark.usefixtures('cpu_double')


def test_weighted_sample_quantile():
    """tests/unit/test_utils.py:64-75."""
    from elfi_b200.ops import weighted_sample_quantile
    x = np.arange(11)
    assert weighted_sample_quantile(x, 0.50) == x[5]
    weights = np.array((0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1))
    assert weighted_sample_quantile(x, 0.50, weights=weights) == x[8]
    assert weighted_sample_quantile(x, 3 / 11) == weighted_sample_quantile(np.flip(x).copy(), 3 / 11)


def test_weighted_var():
    """tests/unit/test_utils.py:78-89."""
    from elfi_b200.ops import weighted_var
    std = .3
    x = np.random.RandomState(12345).normal(-2, std, size=1000)
    w = np.array([1] * len(x))
    assert (weighted_var(x, w) - std) < .1
    cov = [[.5, 0], [0, 3.2]]
    x = np.random.RandomState(12345).multivariate_normal([1, 2], cov, size=1000)
    assert np.linalg.norm(weighted_var(x, w) - np.diag(cov)) < .1


def test_gm_distribution_pdf_and_rvs():
    """tests/unit/test_utils.py:92-145."""
    from elfi_b200.samplers import GMDistribution, normalize_weights
    x = [1, 2, -1]
    means = [0, 2]
    weights = normalize_weights([.4, .1])
    d = GMDistribution.pdf(x, means, weights=weights)
    d_true = weights[0] * ss.norm.pdf(x, loc=means[0]) + weights[1] * ss.norm.pdf(x, loc=means[1])
    assert np.allclose(np.asarray(d.cpu() if hasattr(d, 'cpu') else d), d_true)
    x = [[1, 2, -1], [0, 0, 2]]
    means = [[0, 0, 0], [-1, -.2, .1]]
    d = GMDistribution.pdf(x, means, weights=weights)
    d_true = weights[0] * ss.multivariate_normal.pdf(x, mean=means[0]) + \
        weights[1] * ss.multivariate_normal.pdf(x, mean=means[1])
    assert np.allclose(np.asarray(d.cpu() if hasattr(d, 'cpu') else d), d_true)

    means = [[1000, 3], [-1000, -3]]
    N = 10000
    rvs = GMDistribution.rvs(means, weights=[.3, .7], size=N,
                             random_state=np.random.RandomState(12042017))
    rvs = rvs[rvs[:, 0] < 0, :]
    assert np.abs(len(rvs) / N - .7) < .01
    assert np.abs(np.mean(rvs[:, 1]) + 3) < .1
    prior_logpdf = ss.uniform(0, 1).logpdf
    rvs = GMDistribution.rvs([0.8, 0.5], weights=[.3, .7], size=N, prior_logpdf=prior_logpdf)
    assert np.all(np.isfinite(prior_logpdf(rvs)))


def test_numgrad():
    """tests/unit/test_utils.py:148-151."""
    from elfi_b200.samplers import numgrad
    assert np.allclose(numgrad(lambda x: np.log(x), 3), [1 / 3])
    assert np.allclose(numgrad(lambda x: np.prod(x, axis=1), [1, 3, 5]), [15, 5, 3])


def test_model_prior():
    """tests/unit/test_utils.py:154-185."""
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    from elfi_b200.samplers import ModelPrior
    prior = ModelPrior(ma2.get_model(seed_obs=4))
    rv = prior.rvs(size=10)
    assert rv.shape == (10, 2)
    assert np.allclose(prior.pdf(rv), np.exp(prior.logpdf(rv)))
    grads = prior.gradient_logpdf(rv)
    assert grads.shape == rv.shape and np.allclose(grads, 0)
    loc, scale = 2.2, 1.1
    x = np.random.rand()
    node = elfi.Prior('normal', loc, scale, model=elfi.ElfiModel())
    num_grad = ModelPrior(node.model).gradient_logpdf(x)
    assert np.isclose(num_grad, -(x - loc) / scale ** 2, atol=0.01)
