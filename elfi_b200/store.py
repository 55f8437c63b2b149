"""Output pools: keep node outputs per batch so that a later inference can reuse them
(elfi/store.py:17-377 OutputPool / ArrayPool, :380-560 stores, :563-817 NpyArray;
loader side elfi/loader.py:95-129, context side elfi/model/elfi_model.py:126-208).

Same interface and semantics as the reference -- `OutputPool(outputs)`, `ArrayPool`, dictionary-like
stores indexed by batch_index, `save / open / close / flush / delete`, context (batch_size, seed)
checks -- with one difference that is the point of this row: a store keeps
what the node produced.  Summaries and discrepancies computed by the CUDA operators are device
tensors and stay in HBM, so re-running Rejection / SMC with another distance or threshold reads
them at HBM speed instead of re-simulating; `ArrayPool` stores hold such batches in a write-back
cache and spill them to their `.npy` file lazily (on `flush / close / save`, or when the
per-store `resident_limit` is exceeded).  Host arrays go to the file directly.
"""
import logging
import os
import pickle
import shutil

import numpy as np
import numpy.lib.format as npformat

from . import device as dev

logger = logging.getLogger(__name__)

DEFAULT_PREFIX = 'pools'


def _nbytes(value):
    if dev.is_device_array(value):
        return value.numel() * value.element_size()
    return 0


def _host(value):
    return dev.to_host(value) if dev.is_device_array(value) else value


class OutputPool:
    """Node outputs per batch in dictionary-like stores (default: a dict per node)."""

    _pkl_name = '_outputpool.pkl'

    def __init__(self, outputs=None, name=None, prefix=None):
        if outputs is None:
            self.stores = {}
        elif isinstance(outputs, dict):
            self.stores = outputs
        else:
            self.stores = {node: None for node in outputs}
        self.batch_size = None
        self.seed = None
        self.name = name
        self.prefix = prefix or DEFAULT_PREFIX
        if self.path and os.path.exists(self.path):
            raise ValueError("A pool with this name already exists in {}. You can use "
                             "OutputPool.open() to open it.".format(self.prefix))

    # ---- context ----------------------------------------------------------------------------
    @property
    def output_names(self):
        return list(self.stores.keys())

    @property
    def has_context(self):
        return self.seed is not None and self.batch_size is not None

    def set_context(self, context):
        if self.has_context:
            raise ValueError('Context is already set')
        self.batch_size = context.batch_size
        self.seed = context.seed
        if self.name is None:
            self.name = '{}_{}'.format(type(self).__name__.lower(), self.seed)

    # ---- batches ----------------------------------------------------------------------------
    def get_batch(self, batch_index, output_names=None):
        batch = {}
        for node in (output_names or self.output_names):
            store = self.stores[node]
            if store is not None and batch_index in store:
                batch[node] = store[batch_index]
        return batch

    def add_batch(self, batch, batch_index):
        for node, values in batch.items():
            if node not in self.stores:
                continue
            store = self._get_store_for(node)
            if batch_index in store:      # same seed, same batch: the output is the same
                continue
            store[batch_index] = values

    def remove_batch(self, batch_index):
        for store in self.stores.values():
            if store is not None and batch_index in store:
                del store[batch_index]

    def has_store(self, node):
        return node in self.stores

    def get_store(self, node):
        return self.stores[node]

    def add_store(self, node, store=None):
        if self.stores.get(node) is not None:
            raise ValueError("Store for '{}' already exists".format(node))
        self.stores[node] = store if store is not None else self._make_store_for(node)

    def remove_store(self, node):
        return self.stores.pop(node)

    def _get_store_for(self, node):
        if self.stores[node] is None:
            self.stores[node] = self._make_store_for(node)
        return self.stores[node]

    def _make_store_for(self, node):
        return {}

    def __len__(self):
        return max([len(s) for s in self.stores.values() if s is not None], default=0)

    def __getitem__(self, batch_index):
        return self.get_batch(batch_index)

    def __setitem__(self, batch_index, batch):
        return self.add_batch(batch, batch_index)

    def __contains__(self, batch_index):
        return len(self) > batch_index

    def clear(self):
        for store in self.stores.values():
            if store is not None:
                store.clear()

    # ---- device residency ---------------------------------------------------------------------
    def device_bytes(self):
        """Bytes of HBM held by the stores (device tensors in dict stores and write-back caches)."""
        total = 0
        for store in self.stores.values():
            if store is None:
                continue
            if hasattr(store, 'resident_bytes'):
                total += store.resident_bytes()
            elif isinstance(store, dict):
                total += sum(_nbytes(v) for v in store.values())
        return total

    def to_host(self):
        """Move every device-resident batch to host memory (dict stores) / its file (array stores)."""
        for store in self.stores.values():
            if store is None:
                continue
            if hasattr(store, 'flush'):
                store.flush()
            elif isinstance(store, dict):
                for b in list(store):
                    store[b] = _host(store[b])

    # ---- persistence --------------------------------------------------------------------------
    def save(self):
        """Pickle the stores (one file per node) and the pool under `self.path`."""
        if not self.has_context:
            raise ValueError("Pool context is not set, cannot save. Please see the "
                             "set_context method.")
        os.makedirs(self.path, exist_ok=True)
        for node, store in self.stores.items():
            if hasattr(store, 'flush'):
                store.flush()
            payload = store
            if isinstance(store, dict):      # device tensors are saved as host arrays
                payload = {b: _host(v) for b, v in store.items()}
            try:
                with _working_directory(self.path), open(node + '.pkl', 'wb') as f:
                    pickle.dump(payload, f)
            except BaseException:
                raise IOError('Failed to pickle the store for node {}, please check that '
                              'it is pickleable or remove it before saving.'.format(node))
        shell = self.__class__.__new__(self.__class__)
        shell.__dict__.update(self.__dict__)
        shell.stores = {node: None for node in self.stores}
        with open(os.path.join(self.path, self._pkl_name), 'wb') as f:
            pickle.dump(shell, f)

    def close(self):
        """Save and release the stores; the pool is not usable afterwards."""
        self.save()
        for store in self.stores.values():
            if hasattr(store, 'close'):
                store.close()

    def flush(self):
        for store in self.stores.values():
            if hasattr(store, 'flush'):
                store.flush()

    def delete(self):
        for store in self.stores.values():
            if hasattr(store, 'close'):
                store.close()
        if self.path is not None and os.path.exists(self.path):
            shutil.rmtree(self.path)

    @classmethod
    def open(cls, name, prefix=None):
        """Open a pool saved under prefix/name (the folder may have been moved or renamed)."""
        prefix = prefix or DEFAULT_PREFIX
        path = cls._make_path(name, prefix)
        with open(os.path.join(path, cls._pkl_name), 'rb') as f:
            pool = pickle.load(f)
        with _working_directory(path):       # pickled array stores name their files relatively
            for node in list(pool.stores.keys()):
                try:
                    with open(node + '.pkl', 'rb') as f:
                        pool.stores[node] = pickle.load(f)
                except Exception as e:
                    logger.warning('Failed to load the store for node {}. Reason: {}'.format(
                        node, str(e)))
                    del pool.stores[node]
        pool.name = name
        pool.prefix = prefix
        return pool

    @classmethod
    def _make_path(cls, name, prefix):
        return os.path.join(prefix, name)

    @property
    def path(self):
        if self.name is None:
            return None
        return self._make_path(self.name, self.prefix)


class ArrayPool(OutputPool):
    """OutputPool whose default stores are `.npy` files (NpyStore) under prefix/name/."""

    def __init__(self, outputs=None, name=None, prefix=None, resident_limit=None):
        super().__init__(outputs, name, prefix)
        self.resident_limit = resident_limit

    def _make_store_for(self, node):
        if not self.has_context:
            raise ValueError('ArrayPool has no context set')
        os.makedirs(self.path, exist_ok=True)
        return NpyStore(os.path.join(self.path, node), self.batch_size,
                        resident_limit=self.resident_limit)


class _working_directory:
    def __init__(self, path):
        self.path = path

    def __enter__(self):
        self.previous = os.getcwd()
        os.chdir(self.path)

    def __exit__(self, *exc):
        os.chdir(self.previous)


# ------------------------------------------------------------------------------------ stores
class StoreBase:
    """Outputs of one node, a subset of the dictionary interface keyed by batch_index.
    Any dictionary-like object works as a store."""

    def __getitem__(self, batch_index):
        raise NotImplementedError

    def __setitem__(self, batch_index, data):
        raise NotImplementedError

    def __delitem__(self, batch_index):
        raise NotImplementedError

    def __contains__(self, batch_index):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    def clear(self):
        raise NotImplementedError

    def close(self):
        pass

    def flush(self):
        pass


class ArrayStore(StoreBase):
    """Batches as consecutive slices of any array-like: store[i] = array[i*bs:(i+1)*bs].
    Batches can be appended at the end only and removed from the end only."""

    def __init__(self, array, batch_size, n_batches=-1):
        if n_batches == -1:
            if len(array) % batch_size != 0:
                logger.warning("The array length is not divisible by the batch size.")
            n_batches = len(array) // batch_size
        self.array = array
        self.batch_size = batch_size
        self.n_batches = n_batches

    def _rows(self, batch_index):
        start = self.batch_size * batch_index
        return slice(start, start + self.batch_size)

    def __getitem__(self, batch_index):
        return self.array[self._rows(batch_index)]

    def __setitem__(self, batch_index, data):
        if batch_index > self.n_batches:
            raise IndexError("Appending further than to the end of the store array is "
                             "currently not supported.")
        rows = self._rows(batch_index)
        if rows.stop > len(self.array):
            raise IndexError("There is not enough space left in the store array.")
        self.array[rows] = data
        if batch_index == self.n_batches:
            self.n_batches += 1

    def __contains__(self, batch_index):
        return batch_index < self.n_batches

    def __delitem__(self, batch_index):
        if batch_index not in self:
            raise IndexError("Cannot remove, batch index {} is not in the array".format(
                batch_index))
        if batch_index != self.n_batches - 1:
            raise IndexError("Removing batches from the middle of the store array is "
                             "currently not supported.")
        self.n_batches -= 1

    def __len__(self):
        return self.n_batches

    def clear(self):
        if hasattr(self.array, 'clear'):
            self.array.clear()
        self.n_batches = 0

    def flush(self):
        if hasattr(self.array, 'flush'):
            self.array.flush()

    def close(self):
        if hasattr(self.array, 'close'):
            self.array.close()


class NpyStore(ArrayStore):
    """ArrayStore over an appendable `.npy` file with a device write-back cache.

    Device tensors assigned to the store stay in HBM (`resident`, keyed by batch_index) and are
    written to the file in batch order on `flush()` / `close()` or as soon as they exceed
    `resident_limit` bytes; reads prefer the resident copy.  Host arrays are written through."""

    def __init__(self, file, batch_size, n_batches=-1, resident_limit=None):
        array = file if isinstance(file, NpyArray) else NpyArray(file)
        super().__init__(array, batch_size, n_batches)
        self.resident = {}
        self.resident_limit = resident_limit

    def resident_bytes(self):
        return sum(_nbytes(v) for v in self.resident.values())

    def __getitem__(self, batch_index):
        if batch_index in self.resident:
            return self.resident[batch_index]
        return super().__getitem__(batch_index)

    def __setitem__(self, batch_index, data):
        if batch_index > self.n_batches:
            raise IndexError("Appending further than to the end of the store array is "
                             "currently not supported.")
        if dev.is_device_array(data):
            if len(data) != self.batch_size:
                raise ValueError('Batch of length {} does not match the batch size {}'.format(
                    len(data), self.batch_size))
            self.resident[batch_index] = data
            if batch_index == self.n_batches:
                self.n_batches += 1
            if self.resident_limit is not None and self.resident_bytes() > self.resident_limit:
                self.flush()
            return
        self.resident.pop(batch_index, None)
        self._write(batch_index, np.asarray(data))
        if batch_index == self.n_batches:
            self.n_batches += 1

    def _write(self, batch_index, data):
        rows = self._rows(batch_index)
        if rows.start > len(self.array):
            self._spill(upto=batch_index)          # earlier resident batches go first
        if rows.start > len(self.array):
            raise IndexError('Batch {} cannot be written: earlier batches are missing from the '
                             'store array'.format(batch_index))
        if len(data) != self.batch_size:
            raise ValueError('Batch of length {} does not match the batch size {}'.format(
                len(data), self.batch_size))
        if rows.start == len(self.array):
            self.array.append(data)
        else:
            self.array[rows] = data

    def _spill(self, upto=None):
        for b in sorted(self.resident):
            if upto is not None and b >= upto:
                break
            if b in self.resident:
                self._write(b, dev.to_host(self.resident.pop(b)))

    def __delitem__(self, batch_index):
        super().__delitem__(batch_index)
        self.resident.pop(batch_index, None)
        rows = self._rows(batch_index)
        if rows.start < len(self.array):
            self.array.truncate(rows.start)

    def clear(self):
        self.resident.clear()
        self.array.truncate(0)
        self.n_batches = 0

    def flush(self):
        self._spill()
        self.array.flush()

    def close(self):
        self._spill()
        self.array.close()

    def delete(self):
        self.resident.clear()
        self.array.delete()

    def __getstate__(self):
        self._spill()
        state = self.__dict__.copy()
        state['resident'] = {}
        return state


class NpyArray:
    """An appendable `.npy` file (format 2.0) indexed along the first axis.

    The header is written into a fixed-size block (HEADER_BLOCK bytes, space padded), so the
    shape can be rewritten in place after every append; files written by np.save are opened as
    well (their header is re-laid out on the first append if the new shape does not fit).
    The result is always a valid `.npy` file after `flush()` / `close()`."""

    HEADER_BLOCK = 256
    MAX_SHAPE_LEN = None

    def __init__(self, filename, array=None, truncate=False):
        self.filename = filename if filename.endswith('.npy') else filename + '.npy'
        self.fs = None
        self.dtype = None
        self.row_shape = None
        self.length = 0
        self.data_offset = self.HEADER_BLOCK
        self._dirty = False
        if truncate and os.path.exists(self.filename):
            os.remove(self.filename)
        if os.path.exists(self.filename):
            self.fs = open(self.filename, 'r+b')
            self._read_header()
        else:
            self.fs = open(self.filename, 'w+b')
        if array is not None:
            self.append(np.asarray(array))
            self.flush()

    # ---- header -----------------------------------------------------------------------------
    def _read_header(self):
        self.fs.seek(0)
        version = npformat.read_magic(self.fs)
        if version == (1, 0):
            shape, fortran, dtype = npformat.read_array_header_1_0(self.fs)
        else:
            shape, fortran, dtype = npformat.read_array_header_2_0(self.fs)
        if fortran:
            raise ValueError('Fortran-ordered .npy files cannot be appended to')
        self.data_offset = self.fs.tell()
        self._version = version
        self.dtype = dtype
        self.length = shape[0] if shape else 0
        self.row_shape = tuple(shape[1:])

    def _header_bytes(self, block):
        descr = npformat.dtype_to_descr(self.dtype)
        shape = (self.length,) + tuple(self.row_shape)
        text = "{{'descr': {!r}, 'fortran_order': False, 'shape': {!r}, }}".format(descr, shape)
        prefix = npformat.magic(2, 0)
        room = block - len(prefix) - 4 - 1
        if len(text) > room:
            return None
        body = text.encode('latin1') + b' ' * (room - len(text)) + b'\n'
        return prefix + np.uint32(len(body)).tobytes() + body

    def _write_header(self):
        if self.dtype is None:
            return
        head = self._header_bytes(self.data_offset)
        if head is None or getattr(self, '_version', (2, 0)) != (2, 0):
            self._relayout()
            head = self._header_bytes(self.data_offset)
        self.fs.seek(0)
        self.fs.write(head)
        self._version = (2, 0)
        self._dirty = False

    def _relayout(self):
        """Move the data so that it starts after a HEADER_BLOCK-sized (or larger) header."""
        block = self.HEADER_BLOCK
        while self._header_bytes(block) is None:
            block *= 2
        if block == self.data_offset and getattr(self, '_version', (2, 0)) == (2, 0):
            return
        self.fs.seek(self.data_offset)
        payload = self.fs.read(self.length * self._row_bytes())
        self.fs.seek(block)
        self.fs.write(payload)
        self.fs.truncate(block + len(payload))
        self.data_offset = block
        self._version = (2, 0)

    # ---- array interface ----------------------------------------------------------------------
    def _row_bytes(self):
        return int(np.prod(self.row_shape, dtype=np.int64)) * self.dtype.itemsize

    def _check_open(self):
        if self.fs is None or self.fs.closed:
            self.fs = open(self.filename, 'r+b')

    @property
    def shape(self):
        return (self.length,) + tuple(self.row_shape or ())

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def __len__(self):
        return self.length

    def _bounds(self, sl):
        if isinstance(sl, (int, np.integer)):
            i = int(sl) + (self.length if sl < 0 else 0)
            if not 0 <= i < self.length:
                raise IndexError('index {} out of range'.format(sl))
            return i, i + 1, True
        if not isinstance(sl, slice):
            raise IndexError('NpyArray supports integers and slices along the first axis')
        start, stop, step = sl.indices(self.length)
        if step != 1:
            raise IndexError('NpyArray supports contiguous slices only')
        return start, max(stop, start), False

    def __getitem__(self, sl):
        start, stop, scalar = self._bounds(sl)
        if self.dtype is None:
            return np.empty((0,))
        self._check_open()
        self.fs.flush()
        self.fs.seek(self.data_offset + start * self._row_bytes())
        count = (stop - start) * int(np.prod(self.row_shape, dtype=np.int64))
        out = np.fromfile(self.fs, dtype=self.dtype, count=count)
        out = out.reshape((stop - start,) + tuple(self.row_shape))
        return out[0] if scalar else out

    def __setitem__(self, sl, value):
        start, stop, scalar = self._bounds(sl)
        shape = tuple(self.row_shape) if scalar else (stop - start,) + tuple(self.row_shape)
        block = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=self.dtype), shape))
        self._check_open()
        self.fs.seek(self.data_offset + start * self._row_bytes())
        self.fs.write(block.tobytes())

    def append(self, array):
        """Append rows (first axis) to the file; the first append fixes dtype and row shape."""
        array = np.asarray(array)
        if array.ndim == 0:
            array = array.reshape(1)
        if self.dtype is None:
            self.dtype = array.dtype
            self.row_shape = tuple(array.shape[1:])
        elif tuple(array.shape[1:]) != tuple(self.row_shape):
            raise ValueError("Appended array is of different shape: {} vs {}".format(
                tuple(array.shape[1:]), tuple(self.row_shape)))
        elif array.dtype != self.dtype:
            raise ValueError("Appended array is of different dtype: {} vs {}".format(
                array.dtype, self.dtype))
        self._check_open()
        if self._header_bytes(self.data_offset) is None or getattr(self, '_version', (2, 0)) != (2, 0):
            self._relayout()
        self.fs.seek(self.data_offset + self.length * self._row_bytes())
        self.fs.write(np.ascontiguousarray(array).tobytes())
        self.length += len(array)
        self._dirty = True

    def truncate(self, length=0):
        """Keep the first `length` rows."""
        if self.dtype is None:
            return
        self._check_open()
        self.length = min(self.length, int(length))
        self.fs.truncate(self.data_offset + self.length * self._row_bytes())
        self._dirty = True
        self.flush()

    def clear(self):
        self.truncate(0)

    def flush(self):
        if self.fs is None or self.fs.closed:
            return
        if self._dirty:
            self.fs.truncate(self.data_offset + self.length * self._row_bytes())
            self._write_header()
        self.fs.flush()

    def close(self):
        if self.fs is not None and not self.fs.closed:
            self.flush()
            self.fs.close()

    def delete(self):
        """Close and remove the file."""
        if self.fs is not None and not self.fs.closed:
            self.fs.close()
        if os.path.exists(self.filename):
            os.remove(self.filename)
        self.length = 0

    def memmap(self, mode='r'):
        """numpy.memmap over the current contents."""
        self.flush()
        return np.memmap(self.filename, dtype=self.dtype, mode=mode, offset=self.data_offset,
                         shape=self.shape)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # pickled by file name relative to the pool folder (OutputPool.save/open chdir there)
    def __getstate__(self):
        self.flush()
        return {'filename': os.path.basename(self.filename)}

    def __setstate__(self, state):
        self.__init__(os.path.abspath(state['filename']))
