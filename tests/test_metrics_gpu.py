"""Non-Euclidean Distance metrics on the device (elfi_b200_dist_metric_thr_f64): bit-identical to
SciPy's cdist for 'sqeuclidean', 'cityblock', 'chebyshev'; Minkowski to the accuracy of pow;
'seuclidean' (elfi_b200_dist_seuclidean_thr_f64) bit-identical."""
import pytest

import metric_cases as cases

pytestmark = pytest.mark.gpu


def test_operator_matches_scipy():
    cases.case_operator_matches_scipy(exact=True)


def test_distance_nodes_in_a_model():
    cases.case_distance_nodes_in_a_model()


def test_seuclidean_matches_scipy():
    cases.case_seuclidean_matches_scipy(exact=True)


def test_seuclidean_node_in_a_model():
    cases.case_seuclidean_node_in_a_model()
