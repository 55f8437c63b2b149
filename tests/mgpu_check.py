"""Multi-GPU parity check, launched by torchrun (one process per GPU, NCCL):
  * Rejection in quantile mode over W ranks == the single-process golden (same batches);
  * SMC over W ranks: round 0 identical to the golden, later rounds statistically sane;
  * AdaptiveDistanceSMC: moments merged across ranks;
  * throughput mode SMC over W ranks == one rank in groups of W batches, bit for bit;
  * populations are gathered exactly once.
Prints MGPU_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import elfi_b200 as elfi
    from elfi_b200.examples import ma2
    gold = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'ma2_rejection_quantile.npz')))

    m = ma2.get_model(seed_obs=4)
    res = elfi.Rejection(m['d'], batch_size=1000, seed=123).sample(100, quantile=0.01, bar=False)
    per_rank = -(-10 // world)                      # ceil(10 batches / world)
    assert res.n_sim == per_rank * world * 1000, res.n_sim
    if per_rank * world == 10:                      # same batches as the golden single-process run
        assert res.threshold == float(gold['threshold'])
        assert np.array_equal(res.discrepancies, gold['out_d'])
        assert np.array_equal(res.samples['t1'], gold['out_t1'])
        assert np.array_equal(res.samples['t2'], gold['out_t2'])
    # always: identical to ONE rank simulating the same batch indices 0 .. per_rank*world-1
    one = elfi.Rejection(m['d'], batch_size=1000, seed=123, distributed=False).sample(
        100, n_sim=per_rank * world * 1000, bar=False)
    assert np.array_equal(res.discrepancies, one.discrepancies)
    assert np.array_equal(res.samples_array, one.samples_array)

    # every rank must hold the identical result
    chk = torch.tensor([float(res.discrepancies.sum())], dtype=torch.float64, device='cuda')
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    assert all(float(c) == float(chk) for c in allc)

    # SMC: 2 batches in round 0 (rank 0 -> batch 0, rank 1 -> batch 1) == single-process run
    single = elfi.SMC(m['d'], batch_size=200, seed=123, distributed=False).sample(
        200, quantiles=[.5, .5, .5], bar=False)
    smc = elfi.SMC(m['d'], batch_size=200, seed=123).sample(200, quantiles=[.5, .5, .5], bar=False)
    if world == 2:
        for k in ('d', 't1', 't2'):
            assert np.array_equal(smc.populations[0].outputs[k], single.populations[0].outputs[k]), k
    means = smc.sample_means_array
    assert abs(means[0] - 0.6) < 0.2 and abs(means[1] - 0.2) < 0.2, means
    assert np.all(np.isfinite(smc.weights)) and smc.weights.min() > 0
    thr = [p.threshold for p in smc.populations]
    assert thr[0] > thr[1] > thr[2]
    w_chk = torch.tensor([float(smc.weights.sum())], dtype=torch.float64, device='cuda')
    allw = [torch.zeros_like(w_chk) for _ in range(world)]
    dist.all_gather(allw, w_chk)
    assert all(float(c) == float(w_chk) for c in allw), 'ranks disagree on the SMC population'

    # adaptive distance: 2 batches in round 0; column moments Chan-merged across ranks
    def adaptive(distributed):
        m2 = ma2.get_model(seed_obs=4)
        m2['d'].become(elfi.AdaptiveDistance(m2['S1'], m2['S2']))
        return elfi.AdaptiveDistanceSMC(m2['d'], batch_size=100, seed=11,
                                        distributed=distributed).sample(100, rounds=2, quantile=0.5,
                                                                        bar=False)
    ad1, ad = adaptive(False), adaptive(True)
    if world == 2:
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

    # throughput mode (device priors / simulator / proposals keyed by the global batch index):
    # W ranks == ONE rank processing the same batches in groups of W, BIT FOR BIT, every
    # population: thresholds, particles, weights (sharded O(N^2) density), covariances, n_sim
    md = ma2.get_device_model(seed_obs=4)

    def run_tp(distributed, n=30000, batch=2500):
        kw = dict(distributed=True) if distributed else dict(distributed=False,
                                                            max_parallel_batches=world)
        return elfi.SMC(md['d'], batch_size=batch, seed=5, device_proposal=ma2.DeviceProposal,
                        **kw).sample(n, quantiles=[.5, .4, .4], bar=False)
    tp, tp1 = run_tp(True), run_tp(False)
    assert tp.n_sim == tp1.n_sim and len(tp.populations) == 3, (tp.n_sim, tp1.n_sim)
    for pa, pb in zip(tp.populations, tp1.populations):
        assert pa.threshold == pb.threshold and pa.n_sim == pb.n_sim
        assert np.array_equal(pa.discrepancies, pb.discrepancies)
        assert np.array_equal(pa.samples_array, pb.samples_array)
        assert np.array_equal(pa.weights, pb.weights), np.abs(pa.weights / pb.weights - 1).max()
        assert np.array_equal(pa.cov, pb.cov)
    assert len(np.unique(tp.discrepancies)) == tp.n_samples
    assert tp.n_sim == sum(p.n_sim for p in tp.populations)

    # a population is gathered once: a second extraction returns the same rows and counters
    rej = elfi.Rejection(m['d'], batch_size=500, seed=7)
    first = rej.sample(50, quantile=0.05, bar=False)
    again = rej.extract_result()
    assert again.n_sim == first.n_sim and np.array_equal(again.discrepancies, first.discrepancies)
    at = elfi.AdaptiveThresholdSMC(m['d'], batch_size=500, seed=2).sample(200, max_iter=3, bar=False)
    assert len(np.unique(at.discrepancies)) == at.n_samples
    assert at.n_sim == sum(p.n_sim for p in at.populations)
    dist.barrier()
    if rank == 0:
        print('MGPU_OK world={}'.format(world), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
