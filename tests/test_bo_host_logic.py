"""CPU tests of the BOLFI / SMC host bookkeeping that never touches the device (mirrors the
reference's tests/unit/test_bo.py:9-133 and the objective arithmetic of samplers.py / bolfi.py)."""
import numpy as np
import pytest

import elfi_oracle as o


def _ma2():
    from elfi_b200.examples import ma2
    return ma2.get_model(seed_obs=4)


def test_lcbsc_beta_and_delta():
    from elfi_b200.bo import GPyRegression, LCBSC
    gp = GPyRegression(['t1', 't2'], bounds={'t1': (-2, 2), 't2': (-1, 1)})
    acq = LCBSC(gp, exploration_rate=10, seed=0)
    for t in (0, 1, 10, 100):
        assert acq._beta(t) == pytest.approx(o.lcbsc_beta(t, 2, 10))
    assert acq.delta == pytest.approx(0.1)
    acq2 = LCBSC(gp, delta=0.25)
    assert acq2.exploration_rate == pytest.approx(4.0)


def test_gp_constructor_argument_checks():
    from elfi_b200.bo import GPyRegression
    with pytest.raises(ValueError):
        GPyRegression('t1', bounds={'t1': (0, 1)})
    with pytest.raises(ValueError):
        GPyRegression(['a', 'b'], bounds={'a': (0, 1)})
    with pytest.raises(ValueError):
        GPyRegression(['a'], bounds=[(0, 1)])
    gp = GPyRegression(['b', 'a'], bounds={'a': (0, 1), 'b': (2, 3)})
    assert gp.bounds == [(2, 3), (0, 1)] and gp.input_dim == 2 and gp.n_evidence == 0
    mu, var = gp.predict(np.zeros((3, 2)))                       # unfitted: prior mean 0, var 1
    assert np.array_equal(mu, np.zeros((3, 1))) and np.array_equal(var, np.ones((3, 1)))
    gm, gv = gp.predictive_gradients(np.zeros((3, 2)))
    assert gm.shape == (3, 2) and not gm.any() and not gv.any()


def test_acquisition_noise_checks_and_bounds():
    from elfi_b200.bo import GPyRegression, LCBSC
    gp = GPyRegression(['t1', 't2'], bounds={'t1': (-2, 2), 't2': (-1, 1)})
    with pytest.raises(ValueError):
        LCBSC(gp, noise_var=-1.0)
    with pytest.raises(ValueError):
        LCBSC(gp, noise_var={'t1': 0.1})
    with pytest.raises(ValueError):
        LCBSC(gp, noise_var=[0.1, 0.1])
    acq = LCBSC(gp, noise_var={'t2': 0.0, 't1': 4.0}, seed=3)
    assert acq.noise_var == [4.0, 0.0]
    x = np.tile([1.9, 0.5], (500, 1))
    xn = acq._add_noise(x.copy())
    assert np.all(xn[:, 0] >= -2) and np.all(xn[:, 0] <= 2) and xn[:, 0].std() > 0.1
    assert np.array_equal(xn[:, 1], x[:, 1])                       # zero variance: untouched


def test_multi_start_minimize():
    from elfi_b200.bo import minimize
    target = np.array([0.3, -0.4])

    def fun(x):
        return float(np.sum((x - target) ** 2))

    def grad(x):
        return 2 * (x - target)
    x, v = minimize(fun, [(-2, 2), (-1, 1)], grad=grad, n_start_points=5,
                    random_state=np.random.RandomState(0))
    np.testing.assert_allclose(x, target, atol=1e-6)
    assert v < 1e-10
    x, v = minimize(fun, [(0.5, 2), (-1, 1)], grad=grad, n_start_points=3,
                    random_state=np.random.RandomState(1))
    assert x[0] == pytest.approx(0.5)                              # clipped to the bounds


def test_bayesian_optimization_bookkeeping():
    import elfi_b200 as elfi
    m = _ma2()
    bo = elfi.BayesianOptimization(m['d'], batch_size=5, initial_evidence=20, update_interval=10,
                                   bounds={'t1': (-2, 2), 't2': (-1, 1)}, seed=1)
    assert bo.n_initial_evidence == 20 and bo.n_precomputed_evidence == 0
    assert bo.acq_batch_size == 5
    bo.set_objective(60)
    assert bo.objective['n_sim'] == 60 and bo._objective_n_batches == 12
    assert [bo._get_acquisition_index(i) for i in (0, 3, 4, 5, 11)] == [-4, -1, 0, 1, 7]
    assert bo.prepare_new_batch(0) is None                        # still initial evidence
    bo2 = elfi.BayesianOptimization(m['d'], batch_size=4, initial_evidence=10,
                                    bounds={'t1': (-2, 2), 't2': (-1, 1)}, seed=1)
    assert bo2.n_initial_evidence == 12                           # rounded up to the batch size
    with pytest.raises(ValueError):
        elfi.BayesianOptimization(m['d'], initial_evidence=-1, bounds={'t1': (-2, 2), 't2': (-1, 1)})
    with pytest.raises(ValueError):
        elfi.BOLFI(m['d'], bounds={'t1': (-2, 2), 't2': (-1, 1)}).fit(None)
    with pytest.raises(ValueError):
        elfi.BOLFI(m['d'], bounds={'t1': (-2, 2), 't2': (-1, 1)}).extract_posterior()


def test_smc_objective_bookkeeping():
    import elfi_b200 as elfi
    m = _ma2()
    smc = elfi.SMC(m['d'], batch_size=1000, seed=1)
    smc.set_objective(200, quantiles=[.5, .25, .1])
    assert smc.objective['round'] == 2
    assert list(smc.objective['thresholds']) == [None, None, None]
    assert smc._rejection.objective['n_batches'] == 1             # ceil(200 / .5 / 1000)
    assert smc.objective['n_batches'] == 1
    smc2 = elfi.SMC(m['d'], batch_size=100, seed=1)
    smc2.set_objective(200, thresholds=[1.0, 0.5])
    assert smc2.objective['round'] == 1 and smc2.current_population_threshold == 1.0
    assert smc2._rejection.objective['threshold'] == 1.0
    # per-round seeds: round 0 uses the seed itself (samplers.py:479-480)
    assert smc2._rejection.seed == 1


def test_adaptive_distance_smc_requirements_and_state():
    import elfi_b200 as elfi
    m = _ma2()
    with pytest.raises(TypeError):
        elfi.AdaptiveDistanceSMC(m['d'], batch_size=100)
    m['d'].become(elfi.AdaptiveDistance(m['S1'], m['S2']))
    ad = elfi.AdaptiveDistanceSMC(m['d'], batch_size=100, seed=2)
    assert ad.output_names == ['d', 't1', 't2', 'S1', 'S2']
    ad.set_objective(50, rounds=3, quantile=0.5)
    assert ad.population_size == 50 and ad._rejection.objective['n_samples'] == 100
    assert ad.objective['round'] == 2 and ad._rejection.adaptive
    node = ad.model['d']
    assert node._s['w'] == [None] and node._s['store'] == [0, 0, 0]
    node._s['scale'] = np.array([2.0, 0.5])
    node.update_distance()
    assert len(node._s['w']) == 2 and np.array_equal(node._s['w'][1], [0.5, 2.0])
    assert node._s['store'] == [0, 0, 0]
    node.init_state()
    assert node._s['w'] == [None]


def test_adaptive_threshold_smc_objective():
    import elfi_b200 as elfi
    m = _ma2()
    ats = elfi.AdaptiveThresholdSMC(m['d'], batch_size=500, seed=2, initial_quantile=0.2)
    ats.set_objective(200, max_iter=4)
    assert ats.objective['round'] == 3 and ats._quantiles[0] == 0.2
    assert ats._rejection.objective['n_batches'] == 2             # ceil(200 / 0.2 / 500)
    from elfi_b200.samplers import DensityRatioEstimation, calculate_densratio_basis_sigma
    assert calculate_densratio_basis_sigma(3.0, 5.0) == pytest.approx(15.0 / 4.0)
    with pytest.raises(NotImplementedError):
        DensityRatioEstimation(optimize=True)
    with pytest.raises(ValueError):
        DensityRatioEstimation().fit(np.zeros((200, 2)), np.zeros((200, 2)))   # sigma missing


def test_lockstep_multi_start_equals_sequential_minimize():
    """minimize_lockstep advances all local optimisations together and evaluates their pending
    points in one batched call; start points, optimiser and selection are those of minimize."""
    from elfi_b200.bo import minimize, minimize_lockstep
    centre = np.array([0.3, -0.4])

    def fun(x):
        x = np.asarray(x)
        return float(np.sum((x - centre) ** 2) + 0.3 * np.sin(5 * x[0]) * np.cos(3 * x[1]))

    def grad(x):
        x = np.asarray(x)
        return 2 * (x - centre) + 0.3 * np.array([5 * np.cos(5 * x[0]) * np.cos(3 * x[1]),
                                                  -3 * np.sin(5 * x[0]) * np.sin(3 * x[1])])
    calls = []

    def batch(X):
        calls.append(len(X))
        return np.array([fun(x) for x in X]), np.array([grad(x) for x in X])
    bounds = [(-2, 2), (-1, 1)]
    x_seq, v_seq = minimize(fun, bounds, grad=grad, n_start_points=7,
                            random_state=np.random.RandomState(4))
    x_par, v_par = minimize_lockstep(batch, bounds, n_start_points=7,
                                     random_state=np.random.RandomState(4))
    assert np.array_equal(x_seq, x_par) and v_seq == v_par
    assert max(calls) == 7 and sum(calls) > 3 * len(calls)          # rounds are shared
    # numerical differentiation inside the optimiser when no gradient is supplied
    x_fd, v_fd = minimize_lockstep(lambda X: (np.array([fun(x) for x in X]), None), bounds,
                                   n_start_points=4, random_state=np.random.RandomState(4),
                                   with_grad=False)
    x_ref, v_ref = minimize(fun, bounds, grad=None, n_start_points=4,
                            random_state=np.random.RandomState(4))
    assert np.array_equal(x_fd, x_ref) and v_fd == v_ref
    # an error in the batched evaluator surfaces in the caller, no thread is left waiting

    def broken(X):
        raise RuntimeError('device lost')
    with pytest.raises(RuntimeError):
        minimize_lockstep(broken, bounds, n_start_points=3, random_state=np.random.RandomState(1))
