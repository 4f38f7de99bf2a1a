#!/usr/bin/env python
"""Crafted batches that isolate the hub walker: (A) one giant segment, (B) many regular hubs.
Run under `ncu --metrics gpu__time_duration.sum` and read the walk_hub2 launch durations."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpnet_b200 import RandomProjectionModule  # noqa: E402

dev = 'cuda:0'
N, B = 1_000_001, 100_000
m = RandomProjectionModule(node_num=N, edge_num=10**9, dim_factor=10, num_layer=3, time_decay_weight=1e-7, device=dev,
                           use_matrix=False, beginning_time=np.float64(0.0), not_scale=False, enforce_dim=-1,
                           decay_mode='lazy', init_p0=False, state_device=dev).to(dev)
m.random_projections[0].data.normal_(0, 0.07)
rng = np.random.default_rng(0)
t = 0.0


def run(src, dst, tag, reps=3):
    global t
    for _ in range(reps):
        ts = np.sort(t + rng.random(len(src)) * 3000.0)
        t = ts[-1]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.update(src, dst, ts)
        e1.record()
        torch.cuda.synchronize()
        print(tag, 'update ms', round(e0.elapsed_time(e1), 4), flush=True)


others = rng.permutation(np.arange(2, N))[:B].astype(np.int64)
run(np.full(B, 1, dtype=np.int64), others, 'A one giant of 100k msgs (+100k singletons)')
hubs = rng.integers(2, 202, B).astype(np.int64)
run(hubs, others, 'B 200 regular hubs of ~500 msgs (+100k singletons)')
g10 = rng.integers(2, 12, B).astype(np.int64)
run(g10, others, 'C 10 giants of ~10k msgs (+100k singletons)')
run(rng.permutation(np.arange(2, N))[:B].astype(np.int64), others, 'D no hubs at all (200k singletons)')
m.check_errors()
