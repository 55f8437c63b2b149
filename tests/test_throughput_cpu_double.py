"""Throughput-mode host logic (device priors / simulators / proposals wired into the ElfiModel
graph and the samplers) on the CPU test double.  The double draws from NumPy's RandomState instead
of the device's Philox streams -- same distributions -- so the statistical assertions of
tests/test_throughput_gpu.py apply unchanged; the stream-level tests (bit-exact fused summaries,
sharding invariance) need the device and stay GPU-only."""
import pytest

import test_throughput_gpu as _gpu

pytestmark = pytest.mark.usefixtures('cpu_double')

test_rejection_throughput_mode_statistics = _gpu.test_rejection_throughput_mode_statistics
test_smc_throughput_mode_statistics = _gpu.test_smc_throughput_mode_statistics
test_gauss_smc_throughput_mode_statistics = _gpu.test_gauss_smc_throughput_mode_statistics
test_gnk_adaptive_distance_smc_throughput_mode_statistics = \
    _gpu.test_gnk_adaptive_distance_smc_throughput_mode_statistics
test_logprior_box_matches_scipy = _gpu.test_logprior_box_matches_scipy
test_logprior_matches_model_prior = _gpu.test_logprior_matches_model_prior
