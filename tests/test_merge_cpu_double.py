"""Sync-free candidate-buffer path on the CPU test double (host logic of ops.CandidateBuffer and
of the device-threshold variant of ops.dist_euclid)."""
import pytest

import merge_cases as cases

pytestmark = pytest.mark.usefixtures('cpu_double')


def test_device_thresholds_and_append():
    cases.case_device_thresholds_and_append()


def test_topn_merge_matches_reference_merge():
    cases.case_topn_merge_matches_reference_merge()
