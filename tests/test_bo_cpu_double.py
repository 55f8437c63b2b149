"""BOLFI / Bayesian-optimisation host logic (elfi_b200/bo.py) on the CPU test double: the GP entry
points are SciPy restatements (tests/abi_double.py), so what is exercised is the surrogate's
bookkeeping, the LCBSC acquisition, the multi-start L-BFGS-B minimiser, the BO loop and
BOLFI.fit / extract_result / posterior -- the bodies are the GPU tests of test_gp_gpu.py."""
import pytest

import test_gp_gpu as _gp

pytestmark = pytest.mark.usefixtures('cpu_double')

test_default_hyperparameters_follow_reference_heuristics = \
    _gp.test_default_hyperparameters_follow_reference_heuristics
test_gp_gradients_match_oracle = _gp.test_gp_gradients_match_oracle
test_lcbsc_value_and_gradient = _gp.test_lcbsc_value_and_gradient
test_bad_pivot_raises = _gp.test_bad_pivot_raises
test_optimize_improves_marginal_likelihood = _gp.test_optimize_improves_marginal_likelihood
test_bolfi_ma2_smoke = _gp.test_bolfi_ma2_smoke


@pytest.mark.parametrize('n,p', [(20, 2), (65, 1)])
def test_gp_predict_wrapper(n, p):
    _gp.test_gp_predict_matches_oracle(n, p)
