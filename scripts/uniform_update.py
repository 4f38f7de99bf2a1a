#!/usr/bin/env python
"""A UNIFORM batch through `update`: B=100,000 edges with endpoints drawn uniformly from
N nodes, d=210, L=3, lazy decay — the regime in which the update really is HBM-bound (no duplicate rows, no hub).
Run under `ncu --set full` to get dram__bytes per launch; on its own it prints the CUDA-event time per call and the
algorithmic GB/s (24*L*d + 24 B per edge)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpnet_b200 import RandomProjectionModule  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nodes', type=int, default=4_000_000)
ap.add_argument('--batch', type=int, default=100_000)
ap.add_argument('--calls', type=int, default=8)
args = ap.parse_args()
dev = 'cuda:0'
N, B, d, L = args.nodes + 1, args.batch, 210, 3
torch.manual_seed(0)
m = RandomProjectionModule(node_num=N, edge_num=10 * N, dim_factor=10, num_layer=L, time_decay_weight=1e-7, device=dev,
                           use_matrix=False, beginning_time=np.float64(0.0), not_scale=False, enforce_dim=-1,
                           decay_mode='lazy').to(dev)
rng = np.random.default_rng(1)
t, ms = 0.0, []
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(args.calls):
    s = torch.from_numpy(rng.integers(1, N, B).astype(np.int64)).to(dev)
    q = torch.from_numpy(rng.integers(1, N, B).astype(np.int64)).to(dev)
    ts = torch.from_numpy(np.sort(t + rng.random(B) * 100.0)).to(dev)
    t = float(ts[-1])
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    m.update(s, q, ts, next_time=t)
    b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
m.check_errors()
alg = B * (24 * L * d + 24)
med = float(np.median(ms[2:]))
print(json.dumps({'workload': f'uniform endpoints, N={N}, B={B}, d={d}, L={L}, lazy', 'update_ms': med,
                  'algorithmic_bytes_per_call': alg, 'algorithmic_GBps': alg / (med * 1e-3) / 1e9}))
