"""The known-answer checks of the reference's tests/unit/test_utils.py on the device
(bodies in tests/test_reference_kats_cpu_double.py)."""
import pytest

import test_reference_kats_cpu_double as _kats

pytestmark = pytest.mark.gpu

test_weighted_sample_quantile = _kats.test_weighted_sample_quantile
test_weighted_var = _kats.test_weighted_var
test_gm_distribution_pdf_and_rvs = _kats.test_gm_distribution_pdf_and_rvs
test_numgrad = _kats.test_numgrad
test_model_prior = _kats.test_model_prior
