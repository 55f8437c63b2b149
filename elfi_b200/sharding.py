"""Pure host logic of the multi-GPU plan (no CUDA): which rank computes what and how per-rank
partial results are merged.  Used by samplers.py; unit-tested on CPU with gloo (world_size 2).

The path shards by rows: batch index b belongs to rank b % world_size
(each batch has its own sub-seed, the reference's own unit of parallelism,
elfi/methods/inference/parameter_inference.py:299-305).  Per population there is one all-gather
of fixed-capacity best-n buffers; adaptive distances add a Chan merge of (n, mean, M2).
"""
from math import ceil

import numpy as np


def batch_index(local_step, rank, world_size):
    """Global batch index of the `local_step`-th batch computed by `rank`."""
    return local_step * world_size + rank


def batches_per_rank(n_batches, world_size):
    """Batches each rank runs so that at least n_batches are computed in total."""
    return int(ceil(n_batches / world_size))


def shard_bounds(n, rank, world_size):
    """Contiguous [lo, hi) slice of n items owned by `rank` (equal capacity ceil(n / size))."""
    per = int(ceil(n / world_size))
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def chan_merge(parts):
    """Merge per-shard (count, mean, M2) column moments with Chan's formula; the result equals
    the single-pass batch update of AdaptiveDistance.add_data (elfi_model.py:1117-1123)."""
    n0, m0, s0 = 0.0, None, None
    for nb, mb, sb in parts:
        if nb == 0:
            continue
        mb = np.asarray(mb, dtype=np.float64)
        sb = np.asarray(sb, dtype=np.float64)
        if m0 is None:
            n0, m0, s0 = float(nb), mb.copy(), sb.copy()
            continue
        n1 = n0 + nb
        delta = mb - m0
        m0 = m0 + delta * (nb / n1)
        s0 = s0 + sb + delta ** 2 * (n0 * nb / n1)
        n0 = n1
    return n0, m0, s0


def merge_topn(keys_per_rank, n):
    """Indices (into the rank-order concatenation) of the n smallest keys: the global best-n
    equals the best-n of the union of the per-rank best-n buffers."""
    allk = np.concatenate(keys_per_rank)
    return np.argsort(allk, kind='stable')[:n]
