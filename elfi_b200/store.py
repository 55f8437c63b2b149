"""Device-resident output pool.

A pool remembers what chosen nodes produced in each batch of an inference so that a later
inference with the same (batch_size, seed) -- another threshold, another distance over the same
summaries, more batches -- reads those outputs back instead of simulating again.  The outputs of
the CUDA operators are device tensors and are kept as they are: the re-run streams them from HBM.
When the pool holds more than `resident_limit` bytes of HBM, the oldest batches are moved to host
memory (lazy spill); a spilled batch is uploaded again by the operator that reads it.

The pool plugs into the batch loop where the reference's does: ``model.execute_batch`` asks
``get_batch`` before running a batch and ``ComputationContext.callback`` hands the finished batch to
``add_batch`` (interfaces of elfi/store.py:17-245, elfi/loader.py:95-129,
elfi/model/elfi_model.py:196-208).  Persistence to disk (the reference's ArrayPool / .npy stores,
elfi/store.py:248-817) is a storage feature outside the hot path and is not provided: a pool that
must outlive the process is `to_host()` + whatever serialisation the caller prefers.
"""
from . import device as dev


def _device_nbytes(value):
    return value.numel() * value.element_size() if dev.is_device_array(value) else 0


class BatchStore(dict):
    """batch_index -> output of one node.  A plain dictionary that knows how much HBM its values
    hold and can move a batch to the host."""

    def device_bytes(self):
        return sum(_device_nbytes(v) for v in self.values())

    def resident_batches(self):
        return sorted(b for b, v in self.items() if dev.is_device_array(v))

    def spill(self, batch_index):
        """Move one batch to host memory; returns the HBM bytes released."""
        value = self[batch_index]
        released = _device_nbytes(value)
        if released:
            self[batch_index] = dev.to_host(value)
        return released


class OutputPool:
    """``OutputPool(['S1', 'S2', 'd'])`` keeps those node outputs, batch by batch.

    `outputs` may also be a dict {node name: store or None} to supply own dictionary-like stores.
    `resident_limit` (bytes, None = unlimited) bounds the HBM held by the pool."""

    def __init__(self, outputs=None, name=None, resident_limit=None):
        if isinstance(outputs, dict):
            self.stores = outputs
        else:
            self.stores = {node: None for node in (outputs or [])}
        self.name = name
        self.resident_limit = resident_limit
        self.batch_size = None
        self.seed = None
        self._resident = 0      # running count of device_bytes(), exact after _recount()

    # ---- the inference this pool belongs to ----------------------------------------------------
    @property
    def has_context(self):
        return self.batch_size is not None and self.seed is not None

    def set_context(self, context):
        """Pin the pool to (batch_size, seed): outputs are only reusable under the same two."""
        if self.has_context:
            raise ValueError('The pool already belongs to batch_size={}, seed={}'.format(
                self.batch_size, self.seed))
        self.batch_size, self.seed = context.batch_size, context.seed
        if self.name is None:
            self.name = 'outputpool_{}'.format(self.seed)

    # ---- stores ----------------------------------------------------------------------------------
    @property
    def output_names(self):
        return list(self.stores)

    def has_store(self, node):
        return node in self.stores

    def get_store(self, node):
        return self.stores[node]

    def add_store(self, node, store=None):
        if self.stores.get(node) is not None:
            raise ValueError("The pool already has a store for '{}'".format(node))
        self.stores[node] = BatchStore() if store is None else store

    def remove_store(self, node):
        store = self.stores.pop(node)
        self._recount()
        return store

    def _live_stores(self):
        return [s for s in self.stores.values() if s is not None]

    # ---- batches ---------------------------------------------------------------------------------
    def get_batch(self, batch_index, output_names=None):
        """{node: stored output} for the nodes that have this batch."""
        found = {}
        for node in (self.stores if output_names is None else output_names):
            store = self.stores[node]
            if store is not None and batch_index in store:
                found[node] = store[batch_index]
        return found

    def add_batch(self, batch, batch_index):
        """Keep the pooled nodes' outputs of a finished batch.  A batch that is already stored is
        left alone: same seed and batch index means same output."""
        for node, value in batch.items():
            if node not in self.stores:
                continue
            if self.stores[node] is None:
                self.stores[node] = BatchStore()
            if batch_index not in self.stores[node]:
                self.stores[node][batch_index] = value
                self._resident += _device_nbytes(value)
        if self.resident_limit is not None and self._resident > self.resident_limit:
            self._enforce_limit()

    def remove_batch(self, batch_index):
        for store in self._live_stores():
            store.pop(batch_index, None)
        self._recount()

    def clear(self):
        for store in self._live_stores():
            store.clear()
        self._recount()

    def __len__(self):
        """Number of batches (of the fullest store)."""
        return max((len(s) for s in self._live_stores()), default=0)

    def __contains__(self, batch_index):
        """Does any store hold this batch?  (Ranks of a distributed run hold every W-th index.)"""
        return any(batch_index in s for s in self._live_stores())

    def __getitem__(self, batch_index):
        return self.get_batch(batch_index)

    def __setitem__(self, batch_index, batch):
        self.add_batch(batch, batch_index)

    # ---- HBM residency ---------------------------------------------------------------------------
    def device_bytes(self):
        """HBM held by the stored outputs (counted afresh: stores may have been edited directly)."""
        return self._recount()

    def _recount(self):
        self._resident = sum(s.device_bytes() if hasattr(s, 'device_bytes')
                             else sum(_device_nbytes(v) for v in s.values())
                             for s in self._live_stores())
        return self._resident

    def _spillable(self):
        return [s for s in self._live_stores() if hasattr(s, 'spill')]

    def _enforce_limit(self):
        """Oldest batches first, all nodes of a batch together, until under the limit.  (The
        running byte count makes the common case -- under the limit -- free of any scan.)"""
        excess = self._recount() - self.resident_limit
        if excess <= 0:
            return
        stores = self._spillable()
        for b in sorted(set(b for s in stores for b in s.resident_batches())):
            for s in stores:
                if b in s:
                    freed = s.spill(b)
                    excess -= freed
                    self._resident -= freed
            if excess <= 0:
                return

    def to_host(self):
        """Move every device-resident batch to host memory."""
        for s in self._live_stores():
            if hasattr(s, 'spill'):
                for b in s.resident_batches():
                    s.spill(b)
            else:
                for b in list(s):
                    if dev.is_device_array(s[b]):
                        s[b] = dev.to_host(s[b])
        self._recount()
