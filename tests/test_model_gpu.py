"""GPU mirror of the reference's node-level unit tests (tests/unit/test_elfi_model.py):
Distance == hand-written Euclidean discrepancy, observed summaries, AdaptiveDistance scale /
nested distances."""
import numpy as np
import pytest
from scipy.spatial.distance import cdist

pytestmark = pytest.mark.gpu


def test_distance_equals_handwritten_discrepancy():
    """tests/unit/test_elfi_model.py:139-153: same RNG state -> array_equal."""
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_model(seed_obs=4)
    out = m.generate(200, ['S1', 'S2', 'd'], seed=9)
    S = np.column_stack([out['S1'].cpu().numpy(), out['S2'].cpu().numpy()])
    obs = np.array([[float(m['S1'].observed.cpu()[0]), float(m['S2'].observed.cpu()[0])]])
    assert np.array_equal(out['d'].cpu().numpy(), cdist(S, obs).ravel())

    def discrepancy(s1, s2, observed):
        return cdist(np.column_stack([s1, s2]), np.column_stack(observed)).ravel()
    elfi.Discrepancy(lambda s1, s2, observed: discrepancy(s1.cpu().numpy(), s2.cpu().numpy(),
                                                          [o.cpu().numpy() for o in observed]),
                     m['S1'], m['S2'], name='d_custom')
    out2 = m.generate(200, ['d', 'd_custom'], seed=9)
    assert np.array_equal(out2['d'].cpu().numpy(), np.asarray(out2['d_custom']))


def test_observed_summaries():
    """tests/unit/test_elfi_model.py:34-45."""
    from elfi_b200.examples import ma2
    m = ma2.get_model(seed_obs=4)
    y = m.observed['MA2']
    assert np.array_equal(m['S1'].observed.cpu().numpy(),
                          np.mean(y[:, 1:] * y[:, :-1], axis=1))


def test_adaptive_distance_scale_and_nested_columns():
    """tests/unit/test_elfi_model.py:185-253: Welford scale == np.std after 10, 20, 21 rows;
    nested distances == sqrt(sum(((sim - obs) / scale)^2))."""
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_model(seed_obs=4)
    m['d'].become(elfi.AdaptiveDistance(m['S1'], m['S2']))
    d = m['d']
    rs = np.random.RandomState(0)
    data = rs.randn(21, 2) * [2.0, 0.3] + [1.0, -1.0]
    d.init_state()
    d.add_data(data[:10, 0], data[:10, 1])
    np.testing.assert_allclose(d._s['scale'], np.std(data[:10], axis=0), rtol=1e-12)
    d.add_data(data[10:20, 0], data[10:20, 1])
    np.testing.assert_allclose(d._s['scale'], np.std(data[:20], axis=0), rtol=1e-12)
    d.add_data(data[20:, 0], data[20:, 1])
    np.testing.assert_allclose(d._s['scale'], np.std(data, axis=0), rtol=1e-12)
    scale = d._s['scale'].copy()
    d.update_distance()
    sim = rs.randn(50, 2)
    obs = np.array([[0.3, -0.2]])
    nd = d.nested_distance(sim, obs).cpu().numpy()
    assert nd.shape == (50, 2)
    assert np.array_equal(nd[:, 0], cdist(sim, obs).ravel())
    assert np.array_equal(nd[:, 1], cdist(sim, obs, 'euclidean', w=(1 / scale) ** 2).ravel())
    np.testing.assert_allclose(nd[:, 1], np.sqrt(np.sum(((sim - obs) / scale) ** 2, axis=1)),
                               rtol=1e-12)
