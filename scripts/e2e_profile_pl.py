#!/usr/bin/env python
"""Host-side breakdown of the end-to-end power-law step (public numpy API) on the GPU box."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tpnet_b200.sharded import ShardedRandomProjection  # noqa: E402
from tpnet_b200.synth import SHAPES  # noqa: E402
import dataclasses

shape = dataclasses.replace(SHAPES['powerlaw'], num_src=int(os.environ.get('PL_NODES', '2000000')))
B = 100_000
dev = torch.device('cuda:0')
steps = bench.powerlaw_steps(shape, B, 40)
m = ShardedRandomProjection(node_num=shape.node_num, edge_num=shape.edge_num, dim_factor=shape.dim_factor,
                            num_layer=shape.num_layer, time_decay_weight=shape.time_decay_weight, device=str(dev),
                            use_matrix=False, beginning_time=np.float64(0.0), not_scale=False, enforce_dim=-1,
                            decay_mode='lazy', ext_rows=1024, p0='device', state_device=dev).to(dev)
m.init_p0_on_device(seed=0)
for s, d, t, _ in steps[:8]:
    m.update(s, d, t)
torch.cuda.synchronize()


def timed(fn, lo=8, n=12):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for st in steps[lo:lo + n]:
        fn(st)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def api(st):
    s, d, t, neg = st
    with torch.no_grad():
        _, pos = m.get_pair_wise_feature(s, d)
        _, ng = m.get_pair_wise_feature(s, neg)
        m.update(s, d, t)
        return float((pos.sum() - ng.sum()).item())


print('api step            ms', timed(api))
print('gram (s,d)          ms', timed(lambda st: m.pair_wise_gram(st[0], st[1])))
print('gram (s,neg)        ms', timed(lambda st: m.pair_wise_gram(st[0], st[3])))
print('update              ms', timed(lambda st: m.update(st[0], st[1], st[2]), lo=20))
x = torch.randn(B, 64, device=dev)
with torch.no_grad():
    print('mlp Bx64            ms', timed(lambda st: m.mlp(x)))
h = m._h
if h.stager is None:
    m.pair_wise_gram(steps[0][0], steps[0][1])
print('stage 2 id arrays   ms', timed(lambda st: h.stager.upload([st[0], st[1]], [1, 1], m.node_num, m._stream())))
print('stage 3 arrays      ms', timed(lambda st: h.stager.upload([st[0], st[1], st[2]], [2, 2, 0], m.node_num, m._stream())))
pr = cProfile.Profile()
pr.enable()
for st in steps[30:40]:
    api(st)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
