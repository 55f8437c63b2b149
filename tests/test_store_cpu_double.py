"""Output pools (elfi_b200/store.py) on the CPU test double, plus the
plain-host parts (NpyArray, ArrayStore) that need no device at all."""
import os

import numpy as np
import pytest

import store_cases as cases


def test_pool_usage(cpu_double):
    cases.case_pool_usage()


def test_array_pool(cpu_double, tmp_path):
    cases.case_array_pool(tmp_path)


def test_pool_restarts(cpu_double, tmp_path):
    cases.case_pool_restarts(tmp_path)


def test_npy_array_is_a_valid_appendable_npy_file(tmp_path):
    from elfi_b200.store import NpyArray
    fn = str(tmp_path / 'a.npy')
    rs = np.random.RandomState(0)
    parts = [rs.rand(5, 3), rs.rand(1, 3), rs.rand(100, 3)]
    arr = NpyArray(fn)
    for i, p in enumerate(parts):
        arr.append(p)
        arr.flush()
        assert np.array_equal(np.load(fn), np.concatenate(parts[:i + 1]))
    full = np.concatenate(parts)
    assert len(arr) == 106 and arr.shape == (106, 3) and arr.size == 318
    assert np.array_equal(arr[5:50], full[5:50]) and np.array_equal(arr[-1], full[-1])
    arr[2:4] = 7.0
    full[2:4] = 7.0
    arr.flush()
    assert np.array_equal(np.load(fn), full) and np.array_equal(arr.memmap(), full)
    with pytest.raises(ValueError):
        arr.append(rs.rand(2, 4))
    with pytest.raises(ValueError):
        arr.append(rs.rand(2, 3).astype(np.float32))
    arr.truncate(10)
    assert np.array_equal(np.load(fn), full[:10])
    arr.close()
    again = NpyArray(fn)                              # reopen and continue
    again.append(full[10:20])
    again.close()
    assert np.array_equal(np.load(fn), full[:20])
    # a file written by np.save (version 1.0 header) can be extended as well
    fn2 = str(tmp_path / 'b.npy')
    np.save(fn2, full[:7])
    ext = NpyArray(fn2)
    assert ext.shape == (7, 3)
    ext.append(full[7:9])
    ext.close()
    assert np.array_equal(np.load(fn2), full[:9])
    ext.delete()
    assert not os.path.exists(fn2)


def test_array_store_rules(tmp_path):
    from elfi_b200.store import ArrayStore, NpyStore
    content = np.random.RandomState(1).rand(40, 2)
    for store in (ArrayStore(content.copy(), 10), NpyStore(str(tmp_path / 's'), 10)):
        if isinstance(store, NpyStore):
            for i in range(4):
                store[i] = content[10 * i:10 * i + 10]
        assert len(store) == 4 and 3 in store and 4 not in store
        assert np.array_equal(store[1], content[10:20])
        batch = np.random.rand(10, 2)
        store[1] = batch
        assert len(store) == 4 and np.array_equal(store[1], batch)
        with pytest.raises(IndexError):
            del store[1]                         # only the last batch can be removed
        with pytest.raises(IndexError):
            store[6] = batch                     # only appending at the end
        del store[3]
        assert len(store) == 3
        store[3] = batch
        assert len(store) == 4
        store.clear()
        assert len(store) == 0
