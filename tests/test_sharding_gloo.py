"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the exchange plan of
SURVEY.md section 8e (one all-gather of fixed-capacity best-n buffers, Chan merge of column moments,
row shards of the O(N^2) density) and compare with the single-process answer."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from elfi_b200 import sharding
    from elfi_b200.samplers import Comm
    comm = Comm()
    assert comm.on and comm.size == world and comm.rank == rank

    # --- batches: rank r computes global batch indices r, r + W, ...
    n_batches = 10
    mine = [sharding.batch_index(i, rank, world) for i in range(sharding.batches_per_rank(n_batches, world))]
    rs_all = [np.random.RandomState(1000 + b) for b in range(n_batches)]
    data = {b: (rs_all[b].rand(500), rs_all[b].randn(500, 2)) for b in range(n_batches)}

    # --- local best-n over my batches, then ONE all-gather and a merge
    n = 40
    d_loc = np.concatenate([data[b][0] for b in mine])
    p_loc = np.vstack([data[b][1] for b in mine])
    o = np.argsort(d_loc, kind='stable')[:n]
    d_g = comm.all_gather_rows(torch.from_numpy(d_loc[o])).numpy()
    p_g = comm.all_gather_rows(torch.from_numpy(p_loc[o])).numpy()
    sel = sharding.merge_topn([d_g], n)
    d_all = np.concatenate([data[b][0] for b in range(n_batches)])
    p_all = np.vstack([data[b][1] for b in range(n_batches)])
    ref = np.argsort(d_all, kind='stable')[:n]
    assert np.array_equal(d_g[sel], d_all[ref])
    assert np.array_equal(p_g[sel], p_all[ref])

    # --- accepted counts: scalar all-reduce keeps the n_batches estimate global
    acc = float(np.sum(d_loc < 0.1))
    assert comm.all_reduce_sum(acc) == float(np.sum(d_all < 0.1))

    # --- adaptive-distance moments: per-rank (n, mean, M2) -> Chan merge == global moments
    S_loc = np.vstack([data[b][1] for b in mine]) * [3.0, 0.5] + [10.0, -4.0]
    pack = np.concatenate([[len(S_loc)], S_loc.mean(0), ((S_loc - S_loc.mean(0)) ** 2).sum(0)])
    allp = comm.all_gather_rows(torch.from_numpy(pack[None, :])).numpy()
    nt, mean, m2 = sharding.chan_merge([(r[0], r[1:3], r[3:5]) for r in allp])
    S_all = p_all * [3.0, 0.5] + [10.0, -4.0]
    assert nt == len(S_all)
    np.testing.assert_allclose(mean, S_all.mean(0), rtol=1e-13)
    np.testing.assert_allclose(np.sqrt(m2 / nt), S_all.std(0), rtol=1e-12)

    # --- density shards: equal-capacity contiguous slices cover [0, N) in rank order
    N = 1001
    lo, hi, per = sharding.shard_bounds(N, rank, world)
    part = torch.full((per,), float('nan'), dtype=torch.float64)
    part[:hi - lo] = torch.arange(lo, hi, dtype=torch.float64)
    full = comm.all_gather_rows(part)[:N]
    assert torch.equal(full, torch.arange(N, dtype=torch.float64))
    open(os.path.join(out_dir, 'ok{}'.format(rank)), 'w').write('ok')
    dist.destroy_process_group()


def test_world_size_2_exchange_plan(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), 'ok{}'.format(r))) for r in range(world))


def test_shard_helpers():
    from elfi_b200 import sharding
    assert [sharding.batch_index(i, 1, 4) for i in range(3)] == [1, 5, 9]
    assert sharding.batches_per_rank(10, 4) == 3
    covered = []
    for r in range(3):
        lo, hi, per = sharding.shard_bounds(10, r, 3)
        covered += list(range(lo, hi))
        assert per == 4
    assert covered == list(range(10))
    assert sharding.shard_bounds(2, 3, 4)[:2] == (2, 2)
