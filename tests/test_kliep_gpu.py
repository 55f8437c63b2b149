"""GPU parity: KLIEP density-ratio estimation and AdaptiveThresholdSMC vs the reference goldens."""
import numpy as np
import pytest

import elfi_oracle as o
from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_kliep_matches_reference_golden():
    from elfi_b200 import ops
    g = load_golden('kliep')
    alpha, max_ratio, steps = ops.kliep_fit(g['x'], g['y'], g['wx'], g['wy'], sigma=float(g['sigma']))
    np.testing.assert_allclose(max_ratio, float(g['max_ratio']), rtol=1e-9)
    theta = g['x'][:100]
    d2 = ((g['x'][:, None, :] - theta[None, :, :]) ** 2).sum(-1)
    ratios = np.exp(-0.5 * d2 / float(g['sigma']) ** 2) @ alpha.cpu().numpy()
    np.testing.assert_allclose(ratios, g['ratios'], rtol=1e-8)


@pytest.mark.parametrize('N,p', [(100, 1), (5000, 2), (20000, 3)])
def test_kliep_matches_oracle(N, p):
    from elfi_b200 import ops
    rs = np.random.RandomState(N)
    x = rs.randn(N, p) * 0.6
    y = rs.randn(N + 37, p) + 0.2
    wx, wy = rs.rand(N) + 0.1, rs.rand(N + 37) + 0.1
    alpha_o, mr_o = o.kliep_fit(x, y, wx, wy, sigma=0.9)
    alpha, mr, _ = ops.kliep_fit(x, y, wx, wy, sigma=0.9)
    np.testing.assert_allclose(alpha.cpu().numpy(), alpha_o, rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(mr, mr_o, rtol=1e-8)
    alpha2, mr2, _ = ops.kliep_fit(x, y, sigma=0.9)
    np.testing.assert_allclose(mr2, o.kliep_fit(x, y, sigma=0.9)[1], rtol=1e-8)


def test_too_few_samples_raises():
    from elfi_b200 import ops
    with pytest.raises(ValueError):
        ops.kliep_fit(np.zeros((10, 2)), np.zeros((10, 2)), sigma=1.0, n_basis=100)


def test_adaptive_threshold_smc_ma2():
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    g = load_golden('ma2_adaptive_threshold_smc')
    m = ma2.get_model(seed_obs=4)
    ats = elfi.AdaptiveThresholdSMC(m['d'], batch_size=500, seed=2)
    res = ats.sample(200, max_iter=4, bar=False)
    assert len(res.populations) == int(g['n_pops'])
    assert res.n_sim == int(g['n_sim'])
    q = np.array([np.nan if v is None else v for v in ats._quantiles], dtype=float)
    np.testing.assert_allclose(q, g['quantiles'], rtol=1e-6)
    for i, pop in enumerate(res.populations):
        pre = 'pop{}_'.format(i)
        np.testing.assert_allclose(pop.threshold, float(g[pre + 'threshold']), rtol=1e-7)
        np.testing.assert_allclose(pop.weights, g[pre + 'weights'], rtol=1e-5)
        for k in ('d', 't1', 't2'):
            np.testing.assert_allclose(pop.outputs[k], g[pre + 'out_' + k], rtol=1e-6, atol=1e-9)
