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


@pytest.mark.parametrize('name,extra', [('MaxVar', {}), ('ExpIntVar', {'d_grid': 0.5}),
                                        ('RandMaxVar', {'n_samples': 20, 'sampler': 'metropolis'}),
                                        ('UniformAcquisition', {})])
def test_bolfi_loop_with_each_acquisition(name, extra):
    """BOLFI.fit driven by each acquisition class (acquisition_method= as in bolfi.py:29-42)."""
    import numpy as np
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    from elfi_b200.samplers import ModelPrior
    m = ma2.get_model(seed_obs=4)
    log_d = elfi.Operation(np.log, m['d'], name='log_d')
    bounds = {'t1': (-2, 2), 't2': (-1, 1)}
    gp = elfi.GPyRegression(m.parameter_names, bounds=bounds)
    kw = dict(model=gp, noise_var=0.1, seed=1, n_inits=3, max_opt_iters=30, **extra)
    if name != 'UniformAcquisition':
        kw['prior'] = ModelPrior(m)
    bolfi = elfi.BOLFI(log_d, batch_size=2, initial_evidence=16, update_interval=8, target_model=gp,
                       acquisition_method=getattr(elfi, name)(**kw), bounds=bounds, seed=1)
    bolfi.fit(n_evidence=24, bar=False)
    assert gp.n_evidence == 24
    x = bolfi.extract_result().x_min
    assert -2 <= x['t1'][0] <= 2 and -1 <= x['t2'][0] <= 1
