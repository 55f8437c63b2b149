"""Non-Euclidean Distance metrics: ops wrapper and Distance nodes on the CPU test double."""
import pytest

import metric_cases as cases

pytestmark = pytest.mark.usefixtures('cpu_double')


def test_operator_matches_scipy():
    cases.case_operator_matches_scipy(exact=True)


def test_distance_nodes_in_a_model():
    cases.case_distance_nodes_in_a_model()


def test_seuclidean_matches_scipy():
    cases.case_seuclidean_matches_scipy(exact=True)


def test_seuclidean_node_in_a_model():
    cases.case_seuclidean_node_in_a_model()
