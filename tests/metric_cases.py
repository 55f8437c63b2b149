"""Bodies of the tests of the non-Euclidean Distance metrics (shared by the CPU-double and GPU
collections): the operator against SciPy's cdist, and Distance nodes inside a model."""
import numpy as np
from scipy.spatial.distance import cdist

CASES = [('sqeuclidean', {}), ('cityblock', {}), ('chebyshev', {}), ('minkowski', {'p': 3.0}),
         ('minkowski', {'p': 1}), ('minkowski', {'p': 2}), ('minkowski', {'p': np.inf})]
SHAPES = [(1000, 128), (333, 16), (64, 17), (500, 3), (200, 130), (1, 40), (0, 8)]


def case_operator_matches_scipy(exact):
    from elfi_b200 import ops
    rs = np.random.RandomState(0)
    for metric, kw in CASES:
        for B, D in SHAPES:
            S, obs = rs.randn(B, D), rs.randn(1, D)
            ref = cdist(S, obs, metric, **kw).ravel() if B else np.empty(0)
            thr = float(np.quantile(ref, 0.3)) if B else 1.0
            d, idx = ops.dist_metric(S, obs, metric, p=kw.get('p', 2.0), threshold=thr)
            d, idx = d.cpu().numpy(), idx.cpu().numpy()
            generic_p = metric == 'minkowski' and kw['p'] not in (1, 2, np.inf)
            if generic_p or not exact:
                np.testing.assert_allclose(d, ref, rtol=1e-14, err_msg=str((metric, kw, B, D)))
            else:
                assert np.array_equal(d, ref), (metric, kw, B, D)
            assert np.array_equal(idx, np.nonzero(d <= thr)[0])
    # strided rows (leading dimension > D) and no threshold
    big = rs.randn(300, 64)
    view = big[:, :32]
    obs = rs.randn(1, 32)
    d, idx = ops.dist_metric(view, obs, 'cityblock')
    assert idx is None and np.array_equal(d.cpu().numpy(), cdist(view, obs, 'cityblock').ravel())


def case_seuclidean_matches_scipy(exact):
    """cdist(..., 'seuclidean', V=V): two running sums + IEEE division, bit for bit."""
    from elfi_b200 import ops
    rs = np.random.RandomState(7)
    for B, D in SHAPES + [(257, 33), (100, 1), (100, 2), (999, 15), (40, 255)]:
        S = rs.randn(B, D) * rs.uniform(0.1, 50)
        obs = rs.randn(1, D)
        V = rs.uniform(0.01, 40, D)
        ref = cdist(S, obs, 'seuclidean', V=V).ravel() if B else np.empty(0)
        thr = float(np.quantile(ref, 0.3)) if B else 1.0
        d, idx = ops.dist_seuclidean(S, obs, V, threshold=thr)
        d, idx = d.cpu().numpy(), idx.cpu().numpy()
        if exact:
            assert np.array_equal(d, ref), (B, D, np.max(np.abs(d - ref)))
        else:
            np.testing.assert_allclose(d, ref, rtol=1e-14)
        assert np.array_equal(idx, np.nonzero(d <= thr)[0])
    big = rs.randn(300, 64)
    view = big[:, 8:40]                   # strided rows, no threshold
    obs, V = rs.randn(1, 32), rs.uniform(0.5, 2, 32)
    d, idx = ops.dist_seuclidean(view, obs, V)
    assert idx is None and np.array_equal(d.cpu().numpy(),
                                          cdist(view, obs, 'seuclidean', V=V).ravel())
    for bad in (np.ones(31), np.ones((2, 32))):
        try:
            ops.dist_seuclidean(view, obs, bad)
        except ValueError:
            continue
        raise AssertionError('expected ValueError for V of shape {}'.format(bad.shape))


def case_seuclidean_node_in_a_model():
    """Distance('seuclidean', S1, S2, V=...) inside the MA2 model == cdist over the stacked
    summaries bit for bit, and the sampler's acceptance path runs on it."""
    import elfi_b200 as elfi
    from elfi_b200 import device as dev
    from elfi_b200.examples import ma2
    m = ma2.get_model(seed_obs=4)
    V = [0.5, 4.0]
    node = elfi.Distance('seuclidean', m['S1'], m['S2'], V=V, name='dv')
    out = m.generate(700, ['S1', 'S2', 'dv'], seed=5)
    S = np.column_stack([dev.to_host(out['S1']), dev.to_host(out['S2'])])
    obs = np.array([float(dev.to_host(m[k].observed).ravel()[0]) for k in ('S1', 'S2')]).reshape(1, 2)
    assert np.array_equal(dev.to_host(out['dv']), cdist(S, obs, 'seuclidean', V=V).ravel())
    res = elfi.Rejection(node, batch_size=1000, seed=3).sample(40, threshold=0.4, bar=False)
    assert res.n_samples == 40 and np.all(res.discrepancies <= 0.4)


def case_distance_nodes_in_a_model():
    """Distance('cityblock' | 'minkowski', p=...) inside the MA2 model == cdist over the stacked
    summaries, and Rejection on such a node accepts exactly the rows under the threshold."""
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_model(seed_obs=4)
    nodes = {'l1': elfi.Distance('cityblock', m['S1'], m['S2'], name='l1'),
             'l3': elfi.Distance('minkowski', m['S1'], m['S2'], p=3, name='l3'),
             'linf': elfi.Distance('chebyshev', m['S1'], m['S2'], name='linf'),
             'sq': elfi.Distance('sqeuclidean', m['S1'], m['S2'], name='sq')}
    out = m.generate(500, ['S1', 'S2'] + list(nodes), seed=11)
    from elfi_b200 import device as dev
    S = np.column_stack([dev.to_host(out['S1']), dev.to_host(out['S2'])])
    obs = np.array([float(dev.to_host(m[k].observed).ravel()[0]) for k in ('S1', 'S2')]).reshape(1, 2)
    for name, (metric, kw) in {'l1': ('cityblock', {}), 'l3': ('minkowski', {'p': 3}),
                               'linf': ('chebyshev', {}), 'sq': ('sqeuclidean', {})}.items():
        got = dev.to_host(out[name])
        np.testing.assert_allclose(got, cdist(S, obs, metric, **kw).ravel(), rtol=1e-14)
    res = elfi.Rejection(m['l1'], batch_size=1000, seed=3).sample(50, threshold=0.3, bar=False)
    assert res.n_samples == 50 and np.all(res.discrepancies <= 0.3)
    ref = elfi.Rejection(m['l1'], batch_size=1000, seed=3).sample(50, n_sim=res.n_sim, bar=False)
    assert np.array_equal(np.sort(ref.discrepancies)[:20], np.sort(res.discrepancies)[:20])
    for bad in (dict(w=[1.0, 2.0]), dict(p=2)):
        node = elfi.Distance('cityblock', m['S1'], m['S2'], **bad)
        try:
            node.generate(3)
        except NotImplementedError:
            continue
        raise AssertionError('expected NotImplementedError for {}'.format(bad))
