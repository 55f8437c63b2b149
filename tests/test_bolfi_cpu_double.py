"""BOLFI posterior, MaxVar-family acquisitions and BOLFI.sample on the CPU test double
(bodies in tests/bolfi_cases.py; reference goldens in tests/golden/bolfi_posterior.npz)."""
import pytest

import bolfi_cases as cases

pytestmark = pytest.mark.usefixtures('cpu_double')


def test_prior_gradient():
    cases.case_prior_gradient()


def test_posterior_matches_reference():
    cases.case_posterior_matches_reference()


def test_maxvar_matches_reference():
    cases.case_maxvar_matches_reference()


def test_expintvar_matches_reference():
    cases.case_expintvar_matches_reference()


def test_other_acquisitions():
    cases.case_other_acquisitions()


def test_bolfi_sample():
    cases.case_bolfi_sample()
