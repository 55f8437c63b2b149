"""Bodies of the output-pool tests, shared by the CPU-double and the GPU collections
(modelled on the reference's tests/functional/test_simulation_reuse.py and
tests/unit/test_store.py:101-167)."""
import os

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


def case_array_pool(tmp_path):
    """ArrayPool: .npy-backed stores with a device write-back cache, save / open / move / delete."""
    elfi, m, sim, s1 = counted_ma2()
    from elfi_b200.store import ArrayPool, ArrayStore, NpyArray, OutputPool
    prefix = str(tmp_path / 'pools')
    pool = ArrayPool(['MA2', 'S1'], prefix=prefix)
    N, bs, total = 50, 100, 1000
    rej_pool = elfi.Rejection(m['d'], batch_size=bs, pool=pool, seed=3)
    means = rej_pool.sample(N, n_sim=total, bar=False).sample_means_array
    assert len(pool.stores['MA2']) == total // bs == len(pool.stores['S1']) == len(pool)
    assert 't1' not in pool.stores
    assert len(pool.stores['MA2'].array) == total        # host simulator output: written through
    assert pool.stores['S1'].resident_bytes() == total * 8 and len(pool.stores['S1'].array) == 0
    batch2 = {k: np.array(np.asarray(v.cpu() if hasattr(v, 'cpu') else v)) for k, v in pool[2].items()}

    pool2 = OutputPool(['MA2', 'S1'])
    elfi.Rejection(m['d'], batch_size=bs, pool=pool2, seed=pool.seed).sample(N, n_sim=total, bar=False)
    for bi in range(total // bs):
        assert np.array_equal(np.asarray(pool.stores['S1'][bi].cpu()),
                              np.asarray(pool2.stores['S1'][bi].cpu()))

    calls = sim.calls
    rej_pool.sample(N, n_sim=total, bar=False)
    rej_new = elfi.Rejection(m['d'], batch_size=bs, pool=pool)
    assert np.array_equal(means, rej_new.sample(N, n_sim=total, bar=False).sample_means_array)
    assert sim.calls == calls

    pool.flush()                                          # lazy spill of the device batches
    assert pool.device_bytes() == 0 and len(pool.stores['S1'].array) == total
    assert np.array_equal(np.load(os.path.join(pool.path, 'S1.npy')),
                          np.concatenate([pool.stores['S1'][b] for b in range(total // bs)]))
    pool.close()
    pool = ArrayPool.open(pool.name, prefix=prefix)
    assert len(pool) == total // bs
    pool.close()
    os.rename(pool.path, pool.path + '_move')
    pool = ArrayPool.open(pool.name + '_move', prefix=prefix)
    assert len(pool) == total // bs
    assert np.array_equal(pool[2]['S1'], batch2['S1']) and np.array_equal(pool[2]['MA2'], batch2['MA2'])
    # an opened pool feeds a new inference without simulating
    calls = sim.calls
    again = elfi.Rejection(m['d'], batch_size=bs, pool=pool).sample(N, n_sim=total, bar=False)
    assert sim.calls == calls and np.array_equal(again.sample_means_array, means)

    r = np.random.rand(3 * bs)
    arr = NpyArray(os.path.join(pool.path, 'test.npy'), r)
    pool.add_store('test', ArrayStore(arr, bs))
    assert len(pool.get_store('test')) == 3 and np.array_equal(pool[2]['test'], r[-bs:])
    pool.delete()
    assert not os.path.exists(pool.path)


def case_pool_restarts(tmp_path):
    """save() then keep appending: a re-opened pool sees the saved batches and continues them."""
    elfi, m, sim, s1 = counted_ma2()
    from elfi_b200.store import ArrayPool
    prefix = str(tmp_path / 'pools')
    pool = ArrayPool(['t1', 'd'], name='test', prefix=prefix)
    rej = elfi.Rejection(m, 'd', batch_size=10, pool=pool, seed=123)
    rej.sample(1, n_sim=30, bar=False)
    pool.save()
    rej = elfi.Rejection(m, 'd', batch_size=10, pool=pool)
    rej.set_objective(3, n_sim=60)
    while not rej.finished:
        rej.iterate()
    pool.get_store('t1').array.fs.flush()       # data reaches the file, the header is not rewritten
    assert len(pool) == 6 and len(pool.stores['t1'].array) == 60

    pool2 = ArrayPool.open('test', prefix=prefix)
    assert len(pool2) == 3 and len(pool2.stores['t1'].array) == 30
    s9pool = elfi.Rejection(m, 'd', batch_size=10, pool=pool2).sample(3, n_sim=90, bar=False)
    pool2.save()
    pool2 = ArrayPool.open('test', prefix=prefix)
    s9loaded = elfi.Rejection(m, 'd', batch_size=10, pool=pool2).sample(3, n_sim=90, bar=False)
    s9 = elfi.Rejection(m, 'd', batch_size=10, seed=123).sample(3, n_sim=90, bar=False)
    for a in (s9pool, s9loaded):
        assert np.array_equal(a.samples['t1'], s9.samples['t1'])
        assert np.array_equal(a.discrepancies, s9.discrepancies)
    pool.delete()
    pool2.delete()
