#!/usr/bin/env python
"""Host-side profile of the end-to-end step (public numpy API) on the GPU box."""
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
from tpnet_b200.synth import SHAPES  # noqa: E402

shape = SHAPES['reddit']
warm, steps = bench.make_steps(shape, 260, seed=0, warm=50)
dev = torch.device('cuda:0')
m = bench.build_module(shape, dev, 'auto', warm[0][2][0])
for s, d, t in warm:
    m.update(s, d, t)
for st in steps[:30]:
    bench.api_step(m, st)
torch.cuda.synchronize()


def timed(fn, n=100):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for st in steps[30:30 + n]:
        fn(st)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print('api_step            us/step', timed(lambda st: bench.api_step(m, st)))


def no_item(st):
    with torch.no_grad():
        m.get_pair_wise_feature(*st['enc_pos']); m.get_pair_wise_feature(*st['enc_neg'])
        m.get_pair_wise_feature(st['src'], st['dst']); m.get_pair_wise_feature(st['src'], st['neg'])
        m.update(st['src'], st['dst'], st['t'])


print('no sum/item         us/step', timed(no_item))


def gram_only(st):
    m.pair_wise_gram(*st['enc_pos']); m.pair_wise_gram(*st['enc_neg'])
    m.pair_wise_gram(st['src'], st['dst']); m.pair_wise_gram(st['src'], st['neg'])
    m.update(st['src'], st['dst'], st['t'])


print('gram only (no mlp)  us/step', timed(gram_only))
print('update only         us/step', timed(lambda st: m.update(st['src'], st['dst'], st['t'])))
print('enc gram only       us/call', timed(lambda st: m.pair_wise_gram(*st['enc_pos'])))
print('dec gram only       us/call', timed(lambda st: m.pair_wise_gram(st['src'], st['dst'])))
x = torch.randn(16000, 64, device=dev)
with torch.no_grad():
    print('mlp 16000x64        us/call', timed(lambda st: m.mlp(x)))

pr = cProfile.Profile()
pr.enable()
for st in steps[130:230]:
    gram_only(st)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
