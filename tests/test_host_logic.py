"""CPU tests of the host-side mirror of the ELFI node / sampler API (no GPU needed): graph
compilation, the reference's deterministic execution order and sub-seeding, host RNG parity of
priors + simulators against goldens produced by the reference, ModelPrior, GMDistribution.rvs."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden


def test_sub_seeds_match_reference():
    from elfi_b200.model import get_sub_seed
    meta = json.load(open(os.path.join(GOLDEN, 'meta.json')))
    assert [int(get_sub_seed(123, i)) for i in range(8)] == meta['sub_seeds_123']
    assert [int(get_sub_seed(1, i)) for i in (10, 100, 1000)] == meta['sub_seeds_1_hi']
    cache = {}
    assert [int(get_sub_seed(123, i, cache=cache)) for i in range(8)] == meta['sub_seeds_123']
    with pytest.raises(ValueError):
        get_sub_seed(np.random.RandomState(0), 0)


def test_host_rng_stream_matches_reference_ma2():
    """priors + simulator consume the per-batch RandomState exactly like elfi/executor.py."""
    from elfi_b200.examples import ma2
    g = load_golden('ma2_generate')
    m = ma2.get_model(seed_obs=4)
    out = m.generate(1000, ['t1', 't2', 'MA2'], seed=123)
    for k in ('t1', 't2', 'MA2'):
        assert np.array_equal(out[k], g[k]), k
    assert np.array_equal(m.observed['MA2'], g['observed_MA2'])


def test_host_rng_stream_matches_reference_gauss():
    from elfi_b200.examples import gauss
    g = load_golden('gauss_generate')
    m = gauss.get_model(n_obs=50, seed_obs=3)
    out = m.generate(500, ['mu', 'sigma', 'gauss'], seed=5)
    for k in ('mu', 'sigma', 'gauss'):
        assert np.array_equal(out[k], g[k]), k


def test_model_prior_logpdf_matches_reference():
    from elfi_b200.examples import gauss, ma2
    from elfi_b200.samplers import ModelPrior
    for name, model in (('ma2_prior_logpdf', ma2.get_model(seed_obs=4)),
                        ('gauss_prior_logpdf', gauss.get_model(n_obs=50, seed_obs=3))):
        g = load_golden(name)
        with np.errstate(divide='ignore'):
            lp = ModelPrior(model).logpdf(g['theta'])
        assert np.array_equal(lp, g['logpdf']), name
    prior = ModelPrior(ma2.get_model(seed_obs=4))
    assert prior.dim == 2 and prior.parameter_names == ['t1', 't2']
    x = prior.rvs(7, random_state=np.random.RandomState(0))
    assert x.shape == (7, 2) and np.all(np.isfinite(prior.logpdf(x)))


def test_compile_adds_observed_twins_and_feeders():
    from elfi_b200 import model as em
    from elfi_b200.examples import ma2
    m = ma2.get_model(seed_obs=1)
    plan = em.compile_plan(m, ['d'])
    for n in ('_MA2_observed', '_S1_observed', '_S2_observed', '_d_observed', '_batch_size',
              '_random_state', 't1', 't2', 'MA2', 'S1', 'S2', 'd'):
        assert plan.has_node(n), n
    assert plan.steps['d'].kwargs['observed'] == '_d_observed'
    assert plan.steps['MA2'].kwargs['random_state'] == '_random_state'
    assert plan.steps['MA2'].kwargs['batch_size'] == '_batch_size'
    assert plan.steps['_S1_observed'].args[0] == '_MA2_observed'
    assert plan.steps['_MA2_observed'].sources == []          # the observed data itself
    assert plan.steps['_d_observed'].args == ['_S1_observed', '_S2_observed']
    assert plan.steps['d'].fuses_accept and not plan.steps['_S1_observed'].fuses_accept
    order = plan.order
    assert sorted(order) == sorted(plan.steps)
    assert order.index('t1') < order.index('t2') < order.index('MA2') < order.index('S1') < \
        order.index('d')
    # pruning: only what the requested outputs depend on survives
    plan2 = em.compile_plan(m, ['t1'])
    assert not plan2.has_node('MA2') and not plan2.has_node('d')
    with pytest.raises(ValueError):
        em.compile_plan(m, ['no_such_node'])


def test_execution_order_is_reference_order():
    """Same order as elfi/executor.py:nx_constant_topological_sort (golden from the reference)."""
    from elfi_b200.model import _constant_topological_order
    g = json.load(open(os.path.join(GOLDEN, 'topo_orders.json')))
    for case in g:
        children = {n: [] for n in case['nodes']}
        for u, v in case['edges']:
            children[u].append(v)
        assert _constant_topological_order(case['nodes'], children.__getitem__) == case['order']
    with pytest.raises(ValueError):
        _constant_topological_order('ab', {'a': 'b', 'b': 'a'}.__getitem__)   # a cycle


def test_node_api_errors_and_become():
    import elfi_b200 as elfi
    m = elfi.ElfiModel()
    with pytest.raises(ValueError):
        elfi.Summary(lambda x: x, model=m, name='s')              # needs a parent
    elfi.Prior('uniform', 0, 1, model=m, name='p')
    with pytest.raises(ValueError):
        elfi.Prior('uniform', 0, 1, model=m, name='p')            # duplicate name
    elfi.Simulator(lambda p, batch_size=1, random_state=None: np.zeros((batch_size, 3)), m['p'],
                   observed=np.zeros((1, 3)), name='sim')
    elfi.Summary(lambda x: x.mean(axis=1), m['sim'], name='s')
    elfi.Distance('euclidean', m['s'], name='d')
    assert m.parameter_names == ['p']
    assert isinstance(m['d'], elfi.Distance) and m['d'].parents[0].name == 's'
    with pytest.raises(ValueError):
        elfi.Distance('seuclidean', m['s'], model=m, name='d2')   # V required
    m0 = elfi.ElfiModel()
    elfi.Constant(1.0, model=m0, name='c')
    elfi.Operation(lambda c: c, m0['c'], name='op')
    with pytest.raises(ValueError):
        elfi.Rejection(m0['op'])                                  # model without parameters
    m['d'].become(elfi.AdaptiveDistance(m['s']))
    assert isinstance(m['d'], elfi.AdaptiveDistance)
    kopy = m.copy()
    assert kopy.name != m.name and kopy.has_node('d')
    assert kopy.record('d') is not m.record('d')
    assert kopy.record('d').attrs is m.record('d').attrs      # shared node state
    assert kopy.get_parents('d') == ['s'] and m.get_children('s') == ['d']
    # hidden constants go with the node that used them; observed data goes with its node
    n_before = len(m.nodes)
    elfi.Prior('uniform', 0, 1, model=m, name='q')
    assert len(m.nodes) == n_before + 3
    m.remove_node('q')
    assert len(m.nodes) == n_before
    m.remove_node('sim')
    assert 'sim' not in m.observed and m.get_parents('s') == []
    with pytest.raises(ValueError):
        m.parameter_names = ['nope']


def test_unsupported_metric_fails_loudly():
    import elfi_b200 as elfi
    m = elfi.ElfiModel()
    elfi.Prior('uniform', 0, 1, model=m, name='p')
    elfi.Simulator(lambda p, batch_size=1, random_state=None: np.zeros((batch_size, 1)), m['p'],
                   observed=np.zeros((1, 1)), name='sim')
    elfi.Summary(lambda x: x[:, 0], m['sim'], name='s')
    elfi.Distance('cosine', m['s'], name='d')
    with pytest.raises(NotImplementedError):
        m.generate(4, ['d'], seed=1)


def test_gm_rvs_matches_reference_stream():
    """GMDistribution.rvs consumes the RandomState like elfi/methods/utils.py:200-261."""
    from elfi_b200.samplers import GMDistribution
    g = load_golden('gm_rvs')
    out = GMDistribution.rvs(g['means'], g['cov'], g['weights'], size=int(g['size']),
                             prior_logpdf=None, random_state=np.random.RandomState(7))
    assert np.array_equal(out, g['draws'])
    from elfi_b200.examples import ma2
    from elfi_b200.samplers import ModelPrior
    prior = ModelPrior(ma2.get_model(seed_obs=4))
    out = GMDistribution.rvs(g['means'], g['cov_wide'], g['weights'], size=int(g['size']),
                             prior_logpdf=prior.logpdf, random_state=np.random.RandomState(8))
    assert np.array_equal(out, g['draws_with_prior'])


def test_rejection_objective_bookkeeping():
    """set_objective arithmetic of samplers.py:100-135 (no batches are run)."""
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_model(seed_obs=4)
    rej = elfi.Rejection(m['d'], batch_size=1000, seed=1)
    rej.set_objective(100)                       # default quantile 0.01
    assert rej.objective['n_batches'] == 10 and rej.objective['threshold'] is None
    rej.set_objective(100, n_sim=2500)
    assert rej.objective['n_batches'] == 3
    rej.set_objective(100, threshold=0.3)
    assert rej.objective['threshold'] == 0.3 and rej.objective['n_batches'] == 1
    smc = elfi.SMC(m['d'], batch_size=1000, seed=1)
    with pytest.raises(ValueError):
        smc.set_objective(10)


def test_gnk_stock_model_matches_reference():
    """g-and-k stock graph (host simulator, identity ss_order, euclidean_multiss): no GPU needed."""
    from elfi_b200.examples import gnk
    g = load_golden('gnk_generate')
    m = gnk.get_model(n_obs=50, seed=7)
    out = m.generate(300, ['A', 'B', 'g', 'k', 'GNK', 'd'], seed=3)
    for k in ('A', 'B', 'g', 'k', 'GNK'):
        assert np.array_equal(out[k], g[k]), k
    assert np.array_equal(np.asarray(out['d']), g['d'])
    assert np.array_equal(m.observed['GNK'], g['observed_GNK'])


def test_result_containers():
    from elfi_b200.results import OptimizationResult, Sample, SmcSample
    outs = {'d': np.array([0.1, 0.2, 0.3]), 't1': np.array([1.0, 2.0, 3.0]),
            't2': np.array([0.0, 1.0, 5.0])}
    s = Sample('Rejection', outs, ['t1', 't2'], discrepancy_name='d', threshold=0.3, n_sim=300,
               accept_rate=0.01, seed=1, n_batches=3)
    assert s.n_samples == 3 and s.dim == 2 and s.threshold == 0.3 and s.n_sim == 300
    assert np.array_equal(s.discrepancies, outs['d'])
    assert s.samples_array.shape == (3, 2) and list(s.samples) == ['t1', 't2']
    assert s.sample_means['t1'] == 2.0 and not s.is_multivariate
    assert getattr(s, '_dev', None) is None
    with pytest.raises(AttributeError):
        s.not_there
    s.weights = np.array([0.0, 0.0, 1.0])
    s.meta['cov'] = np.eye(2)
    assert s.sample_means_array[0] == 3.0 and s.cov.shape == (2, 2)
    with pytest.raises(ValueError):
        SmcSample('SMC', outs, ['t1', 't2'], populations=[s])
    smc = SmcSample('SMC', outs, ['t1', 't2'], populations=[s, s], weights=np.ones(3), threshold=0.1)
    assert smc.n_populations == 2 and smc.threshold == 0.1
    opt = OptimizationResult(x_min={'t1': 0.5}, method_name='BO', outputs=outs,
                             parameter_names=['t1', 't2'], n_sim=5)
    assert opt.x_min['t1'] == 0.5 and opt.n_sim == 5
    import copy
    assert copy.copy(s).threshold == 0.3


def test_become_drops_dangling_inputs_and_errors_name_the_node():
    """A node that vanishes in become() leaves no dangling parent names behind, and an exception
    raised inside an operation is re-raised naming the node -- also for exception types whose
    constructor does not take a single message."""
    import elfi_b200 as elfi

    class Picky(Exception):
        def __init__(self, a, b):
            super().__init__(a, b)

    def boom(x):
        raise Picky(1, 2)

    m = elfi.ElfiModel()
    elfi.Constant(1.0, model=m, name='c')
    elfi.Operation(lambda c: c + 1, m['c'], name='a')
    elfi.Operation(lambda c: c + 2, m['c'], name='b')
    elfi.Operation(lambda a, b: a + b, m['a'], m['b'], name='s')
    elfi.Operation(lambda b: b, m['b'], name='user_of_b')
    assert m.generate(1, ['s'])['s'] == 5.0
    m['a'].become(m['b'])                       # 'a' takes b's operation; 'b' disappears
    assert not m.has_node('b') and m.get_parents('s') == ['a'] and m.get_parents('user_of_b') == []
    assert m.generate(1, ['a'])['a'] == 3.0
    elfi.Operation(boom, m['c'], name='bad')
    with pytest.raises(RuntimeError, match="node 'bad'"):
        m.generate(1, ['bad'])
    elfi.Operation(lambda c: 1 / 0, m['c'], name='div')
    with pytest.raises(ZeroDivisionError, match="node 'div'"):
        m.generate(1, ['div'])


def test_observed_side_must_be_deterministic_and_twins_follow_observable_parents():
    """compile_plan: a discrepancy's observed twin may only depend on deterministic nodes
    (elfi/compiler.py:83-104 raises the same way); an Operation between a simulator and a summary
    is not observable, so the summary's twin would re-run it on simulated data -> rejected."""
    import elfi_b200 as elfi
    from elfi_b200 import model as em

    def sim(p, batch_size=1, random_state=None):
        return random_state.normal(p[:, None] if np.ndim(p) else p, 1.0, size=(batch_size, 4))

    m = elfi.ElfiModel()
    elfi.Prior('uniform', 0, 1, model=m, name='p')
    elfi.Simulator(sim, m['p'], observed=np.ones((1, 4)), name='y')
    elfi.Summary(lambda y: y.mean(axis=1), m['y'], name='s')
    elfi.Discrepancy(lambda s, observed: np.abs(s - observed[0]), m['s'], name='d')
    plan = em.compile_plan(m, ['d'])
    assert plan.steps['_s_observed'].args == ['_y_observed']
    assert plan.steps['_d_observed'].args == ['_s_observed']
    # twins carry no per-batch inputs: the simulator's twin is the observed data itself
    assert plan.steps['_y_observed'].kwargs == {} and plan.steps['y'].kwargs.keys() == \
        {'batch_size', 'random_state'}
    out = m.generate(5, ['d', 's'], seed=3)
    assert out['d'].shape == (5,) and np.array_equal(out['d'], np.abs(out['s'] - 1.0))
    assert np.array_equal(m['s'].observed, [1.0])
    # a prior feeding a discrepancy directly: its observed side would be random
    elfi.Discrepancy(lambda p, observed: p, m['p'], name='bad')
    with pytest.raises(ValueError, match='deterministic'):
        em.compile_plan(m, ['bad'])
    m.remove_node('bad')
    # uses_meta: the batch's meta dict reaches the operation
    seen = {}
    elfi.Operation(lambda s, meta=None: seen.update(meta) or s, m['s'], name='tap')
    m['tap'].uses_meta = True
    m.generate(2, ['tap'], seed=9)
    assert seen['batch_index'] == 0 and seen['master_seed'] == 9 and seen['model_name'] == m.name
    assert em.compile_plan(m, ['tap']).has_node('_meta') and not plan.has_node('_meta')
