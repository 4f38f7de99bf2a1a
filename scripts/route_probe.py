#!/usr/bin/env python
"""World-8 routed pair-wise call with all ranks in this process (LocalPeerGroup) on ONE GPU: per-kernel durations of
the routing front end, the pull and the rank-local kernels for a job batch of 800,000 pairs (run under
`ncu --metrics gpu__time_duration.sum`; the pull here goes over local HBM, not NVLink)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_sharded import _sim_ranks, _each, _update_all  # noqa: E402

world = int(os.environ.get('PROBE_WORLD', 8))
N = int(os.environ.get('PROBE_NODES', 2_000_001))
B = int(os.environ.get('PROBE_BATCH', 100_000)) * world
kw = dict(node_num=N, edge_num=10**9, dim_factor=10, num_layer=3, time_decay_weight=1e-7, device='cuda:0',
          use_matrix=False, beginning_time=np.float64(0.0), not_scale=False, enforce_dim=210, p0='device')
ranks, streams = _sim_ranks(world, kw, 'lazy', ext_rows=3 * (B * 35 // (10 * world)) + 4096)
for m in ranks:
    m.init_p0_on_device(seed=0)
rng = np.random.default_rng(0)
t = 0.0
for rep in range(3):
    s = (1 + (rng.zipf(1.2, B) - 1) % (N - 1)).astype(np.int64)
    d = (1 + (rng.zipf(1.2, B) - 1) % (N - 1)).astype(np.int64)
    ts = np.sort(t + rng.random(B) * 3000.0)
    t = float(ts[-1])
    ds, dd, dt = (torch.from_numpy(x).to('cuda:0') for x in (s, d, ts))
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.no_grad():
        ev[0].record(streams[0])
        with torch.cuda.stream(streams[0]):
            ranks[0].routed_pair_wise_feature(ds, dd)
        ev[1].record(streams[0])
        torch.cuda.synchronize()
        print('rank 0 routed pair-wise call of %d job pairs: %.3f ms' % (B, ev[0].elapsed_time(ev[1])), flush=True)
        _each(ranks[1:], streams[1:], lambda m: m.routed_pair_wise_feature(ds, dd))
    _update_all(ranks, streams, ds, dd, dt, next_time=t)
for m in ranks:
    m.check_errors()
