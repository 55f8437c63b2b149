"""The N > 1 sampler path end to end on CPU: world_size-2 gloo processes, C ABI replaced by the
CPU test double (tests/abi_double.py).  Same assertions as tests/mgpu_check.py makes under
torchrun + NCCL on the GPUs: batch index b runs on rank b % W, one all-gather of the best-n
buffers per population, Chan-merged AdaptiveDistance moments, sharded mixture density --
and the result equals a single rank simulating the same batches."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import abi_double
    patch = pytest.MonkeyPatch()
    abi_double.install(patch)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    gold = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'ma2_rejection_quantile.npz')))

    # Rejection, quantile mode: 10 batches over 2 ranks == the reference's single-process golden
    m = ma2.get_model(seed_obs=4)
    res = elfi.Rejection(m['d'], batch_size=1000, seed=123).sample(100, quantile=0.01, bar=False)
    assert res.n_sim == 10000
    assert res.threshold == float(gold['threshold'])
    assert np.array_equal(res.discrepancies, gold['out_d'])
    assert np.array_equal(res.samples['t1'], gold['out_t1'])
    assert np.array_equal(res.samples['t2'], gold['out_t2'])
    # threshold mode stops on a global accepted count; every rank ends with the same sample
    thr = elfi.Rejection(m['d'], batch_size=500, seed=7).sample(50, threshold=0.2, bar=False)
    one = elfi.Rejection(m['d'], batch_size=500, seed=7, distributed=False).sample(
        50, n_sim=thr.n_sim, bar=False)
    assert np.all(thr.discrepancies <= 0.2) and thr.n_samples == 50
    assert np.array_equal(np.sort(one.discrepancies)[:10], np.sort(thr.discrepancies)[:10])

    # SMC: round 0 is rank 0 -> batch 0, rank 1 -> batch 1 == single-process round 0
    single = elfi.SMC(m['d'], batch_size=200, seed=123, distributed=False).sample(
        200, quantiles=[.5, .5, .5], bar=False)
    smc = elfi.SMC(m['d'], batch_size=200, seed=123).sample(200, quantiles=[.5, .5, .5], bar=False)
    for k in ('d', 't1', 't2'):
        assert np.array_equal(smc.populations[0].outputs[k], single.populations[0].outputs[k]), k
    means = smc.sample_means_array
    assert abs(means[0] - 0.6) < 0.25 and abs(means[1] - 0.2) < 0.25, means
    assert np.all(np.isfinite(smc.weights)) and smc.weights.min() > 0
    t = [p.threshold for p in smc.populations]
    assert t[0] > t[1] > t[2]

    # AdaptiveDistanceSMC: per-rank column moments are Chan-merged before the weights update
    def adaptive(distributed):
        m2 = ma2.get_model(seed_obs=4)
        m2['d'].become(elfi.AdaptiveDistance(m2['S1'], m2['S2']))
        return elfi.AdaptiveDistanceSMC(m2['d'], batch_size=100, seed=11, distributed=distributed
                                        ).sample(100, rounds=2, quantile=0.5, bar=False)
    ad1, ad = adaptive(False), adaptive(True)
    np.testing.assert_allclose(ad.populations[0].adaptive_distance_w,
                               ad1.populations[0].adaptive_distance_w, rtol=1e-10)
    for k in ('d', 't1', 't2'):
        np.testing.assert_allclose(ad.populations[0].outputs[k], ad1.populations[0].outputs[k],
                                   rtol=1e-9)
    # the summary columns are not exchanged: each rank keeps those of the population rows it owns
    pop = ad.populations[0]
    rows = pop.local_rows.cpu().numpy()
    assert 'S1' not in pop.outputs and 0 < len(rows) < pop.n_samples
    for k in ('S1', 'S2'):
        np.testing.assert_allclose(pop.local_summaries[k].cpu().numpy(),
                                   ad1.populations[0].outputs[k][rows], rtol=1e-9)
    # ... unless they are asked for by name
    m3 = ma2.get_model(seed_obs=4)
    m3['d'].become(elfi.AdaptiveDistance(m3['S1'], m3['S2']))
    named = elfi.AdaptiveDistanceSMC(m3['d'], output_names=['S1'], batch_size=100, seed=11).sample(
        100, rounds=2, quantile=0.5, bar=False)
    np.testing.assert_allclose(named.populations[0].outputs['S1'], ad1.populations[0].outputs['S1'],
                               rtol=1e-9)
    np.testing.assert_allclose(named.weights, ad.weights, rtol=1e-9)
    assert np.all(np.isfinite(ad.weights)) and len(ad.populations) == 2

    # AdaptiveThresholdSMC: every rank fits the same density ratio and picks the same quantiles
    ats = elfi.AdaptiveThresholdSMC(m['d'], batch_size=500, seed=2)
    at = ats.sample(200, max_iter=4, bar=False)
    q = np.array([np.nan if v is None else v for v in ats._quantiles], dtype=float)
    assert len(at.populations) >= 2 and np.all(np.isfinite(at.weights))
    assert q[0] == 0.2 and np.all((q[1:len(at.populations)] >= 0.05) & (q[1:len(at.populations)] <= 1))
    t = [p.threshold for p in at.populations]
    assert all(a > b for a, b in zip(t, t[1:]))
    # the last population is gathered exactly once (extract_result after update used to gather
    # the already global buffers again: duplicated rows and n_sim multiplied by the world size)
    assert len(np.unique(at.discrepancies)) == at.n_samples
    assert at.n_sim == sum(p.n_sim for p in at.populations)
    assert len(np.unique(smc.discrepancies)) == smc.n_samples
    rej = elfi.Rejection(m['d'], batch_size=500, seed=7)
    first = rej.sample(50, quantile=0.05, bar=False)
    again = rej.extract_result()
    assert again.n_sim == first.n_sim == 1000
    assert np.array_equal(again.discrepancies, first.discrepancies)

    # throughput mode (device priors / simulator / proposals keyed by the global batch index):
    # W ranks == ONE rank that processes its batches in groups of W, bit for bit, every
    # population (thresholds, particles, weights, covariances, n_sim)
    md = ma2.get_device_model(seed_obs=4)

    def run_tp(distributed):
        kw = dict(distributed=True) if distributed else dict(distributed=False,
                                                            max_parallel_batches=world)
        return elfi.SMC(md['d'], batch_size=250, seed=5, device_proposal=ma2.DeviceProposal,
                        **kw).sample(600, quantiles=[.5, .4, .4], bar=False)
    tp, tp1 = run_tp(True), run_tp(False)
    assert tp.n_sim == tp1.n_sim and len(tp.populations) == 3
    for pa, pb in zip(tp.populations, tp1.populations):
        assert pa.threshold == pb.threshold and pa.n_sim == pb.n_sim
        assert np.array_equal(pa.discrepancies, pb.discrepancies)
        assert np.array_equal(pa.samples_array, pb.samples_array)
        assert np.array_equal(pa.weights, pb.weights)
        assert np.array_equal(pa.cov, pb.cov)

    # a rank-local OutputPool in a distributed run: rank r stores the batches it ran (indices
    # r, r + W, ...) and a second inference over the same pool reads them back -- same result,
    # no simulation (the file-backed ArrayPool of round 1 could not do this and was removed)
    calls = {'n': 0}
    mp_ = ma2.get_model(seed_obs=4)
    sim_op = mp_.record('MA2').op

    def counting(*a, **k):
        calls['n'] += 1
        return sim_op(*a, **k)
    mp_.record('MA2').op = counting
    pool = elfi.OutputPool(['t1', 't2', 'MA2'])
    with_pool = elfi.Rejection(mp_['d'], batch_size=1000, seed=123, pool=pool).sample(
        100, quantile=0.01, bar=False)
    assert np.array_equal(with_pool.discrepancies, gold['out_d'])
    assert sorted(pool.get_store('MA2')) == list(range(rank, 10, world)) and calls['n'] == 5
    assert (rank in pool) and ((1 - rank) not in pool) and len(pool) == 5
    again_p = elfi.Rejection(mp_['d'], batch_size=1000, pool=pool).sample(100, quantile=0.01,
                                                                          bar=False)
    assert calls['n'] == 5 and np.array_equal(again_p.discrepancies, gold['out_d'])

    # all ranks hold identical results
    np.save(os.path.join(out_dir, 'rank{}.npy'.format(rank)),
            np.concatenate([res.discrepancies, smc.weights, smc.samples_array.ravel(), ad.weights,
                            at.weights, at.samples_array.ravel(), np.nan_to_num(q),
                            tp.weights, tp.samples_array.ravel()]))
    dist.barrier()
    dist.destroy_process_group()
    patch.undo()


def test_world_size_2_samplers_on_cpu_double(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (np.load(os.path.join(str(tmp_path), 'rank{}.npy'.format(r))) for r in range(world))
    assert np.array_equal(a, b)
