"""BOLFI posterior / MaxVar / sampling with the GP on the device (rows N3 / N4): the GP moments
and gradients behind every logpdf / gradient evaluation come from elfi_b200_gp_predict_grad_f64.
Same bodies as tests/test_bolfi_cpu_double.py."""
import pytest

import bolfi_cases as cases

pytestmark = pytest.mark.gpu


def test_posterior_matches_reference():
    cases.case_posterior_matches_reference()


def test_maxvar_matches_reference():
    cases.case_maxvar_matches_reference()


def test_expintvar_matches_reference():
    cases.case_expintvar_matches_reference()


def test_lcbsc_acquire_matches_reference():
    cases.case_lcbsc_acquire_matches_reference()


def test_hyper_objective_matches_sklearn():
    cases.case_hyper_objective_matches_sklearn()


def test_incremental_factor_update():
    cases.case_incremental_factor_update()


def test_other_acquisitions():
    cases.case_other_acquisitions()


def test_bolfi_sample():
    cases.case_bolfi_sample()
