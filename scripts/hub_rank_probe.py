#!/usr/bin/env python
"""One GPU playing the hub-owning rank of an N=8 power-law step: ONE target receives `HUB` messages in a single update
(the other endpoints are zipf-distributed).  Times `update` with CUDA events on device-resident ids; run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel durations (walk_hub2_kernel = producers + chains + regular hubs)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpnet_b200 import RandomProjectionModule  # noqa: E402

dev = 'cuda:0'
N = int(os.environ.get('PROBE_NODES', 1_250_001))
HUB = int(os.environ.get('PROBE_HUB', 286_000))
REPS = int(os.environ.get('PROBE_REPS', 4))
m = RandomProjectionModule(node_num=N, edge_num=10**9, dim_factor=10, num_layer=3, time_decay_weight=1e-7, device=dev,
                           use_matrix=False, beginning_time=np.float64(0.0), not_scale=False, enforce_dim=-1,
                           decay_mode='lazy', init_p0=False, state_device=dev).to(dev)
m.random_projections[0].data.normal_(0, 0.07)
rng = np.random.default_rng(0)
t = 0.0
for rep in range(REPS + 2):
    others = (2 + (rng.zipf(1.2, HUB) - 1) % (N - 2)).astype(np.int64)
    src = torch.full((HUB,), 1, dtype=torch.int64, device=dev)
    dst = torch.from_numpy(others).to(dev)
    ts_h = np.sort(t + rng.random(HUB) * 3000.0)
    t = float(ts_h[-1])
    ts = torch.from_numpy(ts_h).to(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.update(src, dst, ts, next_time=t)
    e1.record()
    torch.cuda.synchronize()
    if rep >= 2:
        print('hub of %d messages (+%d to zipf targets): update %.3f ms' % (HUB, HUB, e0.elapsed_time(e1)), flush=True)
m.check_errors()
