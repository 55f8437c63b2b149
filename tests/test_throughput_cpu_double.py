"""Throughput-mode host logic (device priors / simulators / proposals wired into the ElfiModel
graph and the samplers) on the CPU test double, at CPU-friendly sizes.  The double draws from
NumPy's RandomState instead of the device's Philox streams -- same distributions -- so only
statistical assertions apply; the stream-level tests (bit-exact fused summaries, sharding
invariance) and the full-size runs are in tests/test_throughput_gpu.py."""
import numpy as np
import pytest

import test_throughput_gpu as _gpu

pytestmark = pytest.mark.usefixtures('cpu_double')

test_gnk_adaptive_distance_smc_throughput_mode_statistics = \
    _gpu.test_gnk_adaptive_distance_smc_throughput_mode_statistics
test_logprior_box_matches_scipy = _gpu.test_logprior_box_matches_scipy
test_logprior_matches_model_prior = _gpu.test_logprior_matches_model_prior


def test_rejection_on_the_ma2_device_model():
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_device_model(seed_obs=4)
    res = elfi.Rejection(m['d'], batch_size=20000, seed=3).sample(400, quantile=0.01, bar=False)
    assert res.n_sim == 40000
    assert abs(res.sample_means['t1'] - 0.6) < 0.1 and abs(res.sample_means['t2'] - 0.2) < 0.1
    again = elfi.Rejection(m['d'], batch_size=20000, seed=3).sample(400, quantile=0.01, bar=False)
    assert np.array_equal(res.samples_array, again.samples_array)


def test_smc_with_device_proposals_ma2():
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_device_model(seed_obs=4)
    smc = elfi.SMC(m['d'], batch_size=10000, seed=5, device_proposal=ma2.DeviceProposal)
    res = smc.sample(1000, quantiles=[0.2, 0.3, 0.3], bar=False)
    thr = [p.threshold for p in res.populations]
    assert len(thr) == 3 and thr[0] > thr[1] > thr[2]
    means = res.sample_means_array
    assert abs(means[0] - 0.6) < 0.1 and abs(means[1] - 0.2) < 0.1, means
    assert np.all(np.isfinite(res.weights)) and res.weights.min() >= 0


def test_smc_with_device_proposals_gauss():
    import elfi_b200 as elfi
    from elfi_b200.examples import gauss
    m, proposal = gauss.get_device_model(n_obs=50, seed_obs=3)
    res = elfi.SMC(m['d'], batch_size=10000, seed=4, device_proposal=proposal).sample(
        1000, quantiles=[0.2, 0.3, 0.3], bar=False)
    a = res.sample_means_array
    assert abs(a[0] - 4.0) < 0.3 and abs(a[1] - 0.4) < 0.3, a


def test_adaptive_threshold_smc_on_the_gauss_device_model():
    """BASELINE config #3's algorithm (adaptive threshold, KLIEP) on the device Gaussian model."""
    import elfi_b200 as elfi
    from elfi_b200.examples import gauss
    m, proposal = gauss.get_device_model(n_obs=50, seed_obs=3)
    ats = elfi.AdaptiveThresholdSMC(m['d'], batch_size=5000, seed=4, device_proposal=proposal)
    res = ats.sample(500, max_iter=4, bar=False)
    assert 2 <= len(res.populations) <= 4
    thr = [p.threshold for p in res.populations]
    assert all(a > b for a, b in zip(thr, thr[1:]))
    a = res.sample_means_array
    assert abs(a[0] - 4.0) < 0.4 and abs(a[1] - 0.4) < 0.4, a
