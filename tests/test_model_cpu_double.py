"""ElfiModel graph / AdaptiveDistance / AdaptiveThresholdSMC host logic on the CPU test double
(tests/abi_double.py); the bodies are the GPU tests of test_model_gpu.py and test_kliep_gpu.py."""
import pytest

import test_kliep_gpu as _kliep
import test_model_gpu as _model

pytestmark = pytest.mark.usefixtures('cpu_double')

test_distance_equals_handwritten_discrepancy = _model.test_distance_equals_handwritten_discrepancy
test_observed_summaries = _model.test_observed_summaries
test_adaptive_distance_scale_and_nested_columns = _model.test_adaptive_distance_scale_and_nested_columns
test_kliep_matches_reference_golden = _kliep.test_kliep_matches_reference_golden
test_too_few_samples_raises = _kliep.test_too_few_samples_raises
test_adaptive_threshold_smc_ma2 = _kliep.test_adaptive_threshold_smc_ma2
