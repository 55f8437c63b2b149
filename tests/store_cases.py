"""Bodies of the output-pool tests, shared by the CPU-double and the GPU collections
(modelled on the reference's tests/functional/test_simulation_reuse.py and
tests/unit/test_store.py:101-167)."""
import numpy as np
import pytest


class Counter:
    def __init__(self, fn):
        self.fn, self.calls = fn, 0

    def __call__(self, *args, **kwargs):
        self.calls += 1
        return self.fn(*args, **kwargs)


def counted_ma2():
    """MA2 model whose simulator and summary count their invocations."""
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    m = ma2.get_model(seed_obs=4)
    sim = Counter(m.record('MA2').op)
    m.record('MA2').op = sim
    s1 = Counter(m.record('S1').op)
    m.record('S1').op = s1
    return elfi, m, sim, s1


def case_pool_usage():
    """A populated pool replaces the simulator; removing stores recomputes only what is missing."""
    elfi, m, sim, s1 = counted_ma2()
    from elfi_b200 import device as dev
    pool = elfi.OutputPool(outputs=m.parameter_names + ['MA2', 'S1', 'S2', 'd'])
    rej = elfi.Rejection(m['d'], batch_size=500, pool=pool, seed=7)
    res = rej.sample(20, quantile=0.01, bar=False)
    assert rej.pool is pool and pool.has_context and pool.seed == 7 and pool.batch_size == 500
    assert len(pool) == 4 and sim.calls == 4               # observed data is given, not simulated
    first_calls = sim.calls
    # summaries / distances computed by the CUDA operators stay device resident in the pool
    assert dev.is_device_array(pool.get_store('S1')[0]) and dev.is_device_array(pool[0]['d'])
    assert pool.device_bytes() >= 4 * 500 * 8 * 3

    res2 = elfi.Rejection(m['d'], batch_size=500, pool=pool).sample(20, quantile=0.01, bar=False)
    assert sim.calls == first_calls and np.array_equal(res2.discrepancies, res.discrepancies)
    assert np.array_equal(res2.samples_array, res.samples_array)

    pool.remove_store('MA2')             # parameters + discrepancy are enough
    res3 = elfi.Rejection(m['d'], batch_size=500, pool=pool).sample(20, quantile=0.01, bar=False)
    assert sim.calls == first_calls and np.array_equal(res3.discrepancies, res.discrepancies)

    pool.remove_store('d')               # the distance is recomputed from the stored summaries
    s1_calls = s1.calls
    res4 = elfi.Rejection(m['d'], batch_size=500, pool=pool).sample(20, quantile=0.01, bar=False)
    # (the summary runs once per inference on the 1-row observed twin -- its output is cached for
    # the following batches -- and never on simulated data)
    assert sim.calls == first_calls and s1.calls == s1_calls + 1
    assert np.array_equal(res4.discrepancies, res.discrepancies)

    # a different distance over the stored summaries: no simulation, different result
    m['d'].become(elfi.Distance('seuclidean', m['S1'], m['S2'], V=[0.5, 4.0]))
    res5 = elfi.Rejection(m['d'], batch_size=500, pool=pool).sample(20, quantile=0.01, bar=False)
    assert sim.calls == first_calls and not np.array_equal(res5.discrepancies, res.discrepancies)

    # more batches than stored: the missing ones are simulated and added
    res6 = elfi.Rejection(m['d'], batch_size=500, pool=pool).sample(20, n_sim=3000, bar=False)
    assert res6.n_sim == 3000 and len(pool) == 6 and sim.calls == first_calls + 2

    with pytest.raises(ValueError):
        elfi.Rejection(m['d'], batch_size=100, pool=pool)
    with pytest.raises(ValueError):
        elfi.Rejection(m['d'], batch_size=500, pool=pool, seed=8)
    pool.to_host()
    assert pool.device_bytes() == 0 and isinstance(pool.get_store('S1')[0], np.ndarray)


def case_pool_spill():
    """resident_limit: the oldest batches move to host memory, all nodes of a batch together; a
    spilled pool feeds a new inference with the same result."""
    elfi, m, sim, s1 = counted_ma2()
    from elfi_b200 import device as dev
    bs, n_batches = 200, 6
    per_batch = bs * 8 * 3                                    # S1, S2, d: one double per row each
    pool = elfi.OutputPool(['S1', 'S2', 'd'], resident_limit=2 * per_batch)
    res = elfi.Rejection(m['d'], batch_size=bs, pool=pool, seed=11).sample(
        30, n_sim=bs * n_batches, bar=False)
    assert len(pool) == n_batches and pool.device_bytes() == 2 * per_batch
    for node in ('S1', 'S2', 'd'):
        store = pool.get_store(node)
        assert store.resident_batches() == [n_batches - 2, n_batches - 1]
        assert isinstance(store[0], np.ndarray) and dev.is_device_array(store[n_batches - 1])
    calls = sim.calls
    again = elfi.Rejection(m['d'], batch_size=bs, pool=pool).sample(30, n_sim=bs * n_batches,
                                                                    bar=False)
    # parameters are not pooled: they are drawn again (same seed, same values); nothing asks for
    # the simulator since summaries and distance are read back (spilled batches from the host)
    assert np.array_equal(again.discrepancies, res.discrepancies)
    assert np.array_equal(again.samples_array, res.samples_array)
    assert sim.calls == calls
    pool.remove_batch(n_batches - 1)
    assert len(pool) == n_batches - 1 and (n_batches - 1) not in pool and 0 in pool
    pool.clear()
    assert len(pool) == 0 and pool.device_bytes() == 0
    with pytest.raises(ValueError):
        pool.add_store('S1')
    with pytest.raises(ValueError):
        pool.set_context(None)
    own = {}
    pool2 = elfi.OutputPool({'d': own})                        # any dictionary works as a store
    elfi.Rejection(m['d'], batch_size=bs, pool=pool2, seed=11).sample(5, n_sim=bs, bar=False)
    assert list(own) == [0] and pool2.device_bytes() == bs * 8
    pool2.to_host()
    assert isinstance(own[0], np.ndarray) and pool2.device_bytes() == 0
