#!/usr/bin/env python
"""SMC-ABC on MA2 in throughput mode (device priors / simulator / proposals) at scale:
accepted particles per second through the public sampler API, 1..N GPUs (torchrun).

    python scripts/bench_smc.py --population 1000000 --batch 1000000 --pops 5 --quantile 0.5
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/bench_smc.py ...

`--batch` is the per-rank batch size.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--population', dest='n', type=int, default=1_000_000)
    ap.add_argument('--batch', type=int, default=1_000_000)
    ap.add_argument('--pops', type=int, default=5)
    ap.add_argument('--quantile', type=float, default=0.5)
    ap.add_argument('--seed', type=int, default=1)
    ap.add_argument('--warm', type=int, default=1)
    ap.add_argument('--model', default='ma2', choices=['ma2', 'gauss', 'gnk'],
                    help='ma2 (scaling target), gauss (BASELINE config #3) or gnk (config #5: '
                         'AdaptiveDistanceSMC over n_obs order statistics, --pops rounds)')
    ap.add_argument('--n-obs', type=int, default=256, help='observations per simulation (gnk)')
    ap.add_argument('--adaptive-threshold', action='store_true',
                    help='AdaptiveThresholdSMC (KLIEP quantile selection, BASELINE config #3) '
                         'instead of fixed quantiles; --pops is then max_iter')
    args = ap.parse_args()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import elfi_b200 as elfi
    from elfi_b200.examples import gauss, gnk, ma2

    if args.model == 'ma2':
        m, proposal = ma2.get_device_model(seed_obs=4), ma2.DeviceProposal
    elif args.model == 'gauss':
        m, proposal = gauss.get_device_model(n_obs=50, seed_obs=3)
    else:
        m, proposal = gnk.get_device_model(n_obs=args.n_obs, seed=7)

    def run(n, batch, pops):
        if args.model == 'gnk':
            smc = elfi.AdaptiveDistanceSMC(m['d'], batch_size=batch, seed=args.seed,
                                           device_proposal=proposal)
            return smc.sample(n, rounds=pops, quantile=args.quantile, bar=False)
        if args.adaptive_threshold:
            smc = elfi.AdaptiveThresholdSMC(m['d'], batch_size=batch, seed=args.seed,
                                            device_proposal=proposal)
            return smc.sample(n, max_iter=pops, bar=False)
        smc = elfi.SMC(m['d'], batch_size=batch, seed=args.seed, device_proposal=proposal)
        return smc.sample(n, quantiles=[args.quantile] * pops, bar=False)

    for _ in range(args.warm):
        # warm-up at FULL size with two populations (one weights step): contexts, scratch arenas,
        # the caching allocator and the NCCL buffers reach their final sizes before the timed run
        run(args.n, args.batch, 2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from elfi_b200.samplers import COMM_STATS, PHASES
    PHASES.tot.clear()
    COMM_STATS.update(all_gather_calls=0, all_gather_bytes=0)
    t0 = time.perf_counter()
    res = run(args.n, args.batch, args.pops)
    final = {k: res.outputs[k] for k in [res.discrepancy_name] + list(res.parameter_names)}
    w_final = res.weights
    torch.cuda.synchronize()
    my_dt = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    per_rank = [my_dt]
    if world > 1:
        allv = torch.empty(world, dtype=torch.float64, device='cuda')
        dist.all_gather_into_tensor(allv, torch.tensor([my_dt], dtype=torch.float64, device='cuda'))
        per_rank = [round(v, 4) for v in allv.tolist()]
    if int(os.environ.get('RANK', '0')) == 0:
        accepted = args.n * len(res.populations)
        out = {'bench': 'smc_abc_{}_throughput_mode'.format(args.model), 'n_gpus': world, 'population': args.n,
               'populations': len(res.populations), 'adaptive_threshold': bool(args.adaptive_threshold), 'quantile': args.quantile, 'batch_per_rank': args.batch,
               'seconds': dt, 'accepted_particles_per_s': accepted / dt,
               'simulated': int(res.n_sim), 'simulated_per_s': res.n_sim / dt,
               'pair_terms': float(args.n) ** 2 * (len(res.populations) - 1),
               'pair_terms_per_s': float(args.n) ** 2 * (len(res.populations) - 1) / dt,
               'posterior_means': [float(v) for v in res.sample_means_array],
               'thresholds': [float(p.threshold) for p in res.populations],
               'per_rank_seconds': per_rank,
               'all_gather_calls': COMM_STATS['all_gather_calls'],
               'all_gather_bytes_received_per_rank': COMM_STATS['all_gather_bytes'],
               'exchange_MB_per_generation': COMM_STATS['all_gather_bytes'] / 1e6 /
               max(1, len(res.populations)),
               'n_sim_per_population': [int(p.n_sim) for p in res.populations]}
        if hasattr(res.populations[-1], 'adaptive_distance_w'):
            w_ad = res.populations[-1].adaptive_distance_w
            out['adaptive_distance_w_range'] = [float(min(w_ad)), float(max(w_ad))]
        from elfi_b200.samplers import PHASES
        if PHASES.on:
            out['phases_s'] = PHASES.report()
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
