#!/usr/bin/env python
"""BASELINE.json configs[4]: projection-update micro-benchmark sweep on one GPU.

    python scripts/sweep_update.py [--quick] > gpurun_out/sweep.jsonl

N = 1,000,001 nodes, d in {64..1024}, L in {1..4}, batch in {200..100k}, uniform vs zipf(1.5) targets
(SURVEY.md 8(d) item 5).  Per configuration: device-resident ids, `update` and the decoder-shaped
pair-wise call (2B pairs) timed with CUDA events, reported as edges/s, pairs/s and algorithmic GB/s
(24*L*d + 24 bytes per edge, 8*(L+1)*d + 4*(2L+2)^2 + 16 bytes per pair)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpnet_b200 import RandomProjectionModule  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--quick', action='store_true', help='a 36-point subset (d 64/256/1024, L 1/3, B 200/20k/100k)')
ap.add_argument('--nodes', type=int, default=1_000_000)
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--dims', type=int, nargs='*')
ap.add_argument('--layers', type=int, nargs='*')
ap.add_argument('--batches', type=int, nargs='*')
args = ap.parse_args()

dev = torch.device('cuda:0')
N = args.nodes + 1
DIMS = [64, 256, 1024] if args.quick else [64, 128, 256, 512, 1024]
LAYERS = [1, 3] if args.quick else [1, 2, 3, 4]
BATCHES = [200, 20_000, 100_000] if args.quick else [200, 2_000, 20_000, 100_000]
DIMS, LAYERS, BATCHES = args.dims or DIMS, args.layers or LAYERS, args.batches or BATCHES
rng = np.random.default_rng(3)
t_start = time.time()


def ids(kind, n):
    if kind == 'uniform':
        return rng.integers(1, N, n).astype(np.int64)
    raw = (rng.zipf(1.5, n) - 1) % (N - 1)
    return (1 + (raw * 2654435761) % (N - 1)).astype(np.int64)      # hubs scattered over the address range


for d in DIMS:
    for L in LAYERS:
        try:
            m = RandomProjectionModule(node_num=N, edge_num=10 ** 9, dim_factor=1, num_layer=L, time_decay_weight=1e-7,
                                       device=str(dev), use_matrix=False, beginning_time=np.float64(0.0),
                                       not_scale=False, enforce_dim=d, decay_mode='lazy', init_p0=False,
                                       state_device=dev).to(dev)
            m.random_projections[0].data.normal_(0, d ** -0.5)
        except Exception as e:                                       # noqa: BLE001
            print(json.dumps({'d': d, 'L': L, 'error': repr(e)}), flush=True)
            continue
        per_edge = 24 * L * d + 24
        per_pair = 8 * (L + 1) * d + 4 * (2 * L + 2) ** 2 + 16
        clock = 0.0
        for B in BATCHES:
            for kind in ('uniform', 'zipf1.5'):
                try:
                    batches = []
                    for _ in range(3 + args.reps):
                        s, t_ = ids(kind, B), ids(kind, B)
                        ts = np.sort(clock + rng.random(B) * 30.0)
                        clock = float(ts[-1])
                        g = lambda a: torch.from_numpy(a).to(dev)     # noqa: E731
                        batches.append((g(s), g(t_), g(ts), clock))
                    for s, t_, ts, last in batches[:3]:
                        m.update(s, t_, ts, next_time=last)
                        m.pair_wise_gram(torch.cat([s, s]), torch.cat([t_, s.flip(0)]))
                    torch.cuda.synchronize()
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.reps)]
                    for i, (s, t_, ts, last) in enumerate(batches[3:]):
                        a, b = torch.cat([s, s]), torch.cat([t_, s.flip(0)])
                        ev[3 * i].record()
                        m.update(s, t_, ts, next_time=last)
                        ev[3 * i + 1].record()
                        m.pair_wise_gram(a, b)
                        ev[3 * i + 2].record()
                    torch.cuda.synchronize()
                    m.check_errors()
                    upd = float(np.median([ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(args.reps)]))
                    prw = float(np.median([ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(args.reps)]))
                    print(json.dumps({'d': d, 'L': L, 'batch': B, 'targets': kind, 'update_ms': upd,
                                      'edges_per_s': B / (upd * 1e-3), 'update_alg_GBps': B * per_edge / (upd * 1e-3) / 1e9,
                                      'pairs': 2 * B, 'pairwise_ms': prw, 'pairs_per_s': 2 * B / (prw * 1e-3),
                                      'pairwise_alg_GBps': 2 * B * per_pair / (prw * 1e-3) / 1e9,
                                      'timing': 'CUDA events around eager calls (includes launch latency), median of %d'
                                                % args.reps}), flush=True)
                except Exception as e:                               # noqa: BLE001
                    print(json.dumps({'d': d, 'L': L, 'batch': B, 'targets': kind, 'error': repr(e)}), flush=True)
        del m
        torch.cuda.empty_cache()
print(json.dumps({'done': True, 'seconds': time.time() - t_start, 'nodes': N}), flush=True)
