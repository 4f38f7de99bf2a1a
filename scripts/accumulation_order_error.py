#!/usr/bin/env python
"""CPU experiment behind `accumulation='chunked'` (include/tpnet_b200.h, tpn_state_t::giant_chunk): how far are
(a) the reference's sequential fp32 order and (b) the chunked order from the EXACT result (the same recurrence in
f64 with the same fp32 weights and decay factors), on a down-scaled replica of the bench workload (power-law,
zipf 1.2, 100,001 nodes, batches of 100,000 x world edges, lambda = 1e-7)?  Per layer: max over rows of
max|row - exact| / max|exact row|.  Uses the oracle only (test infrastructure).

    python scripts/accumulation_order_error.py [--world 1] [--batches 3] > profiles/r02_accumulation_order_error.txt
"""
import argparse
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.walk_projection import WalkProjectionOracle, decay_factors, edge_weights  # noqa: E402
from tpnet_b200.synth import SHAPES, edge_stream  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--world', type=int, default=1)
ap.add_argument('--batches', type=int, default=3)
ap.add_argument('--chunk', type=int, default=1024)
a = ap.parse_args()
shape = dataclasses.replace(SHAPES['powerlaw'], num_src=100_000)
B = 100_000 * a.world
kw = dict(node_num=shape.node_num, edge_num=shape.edge_num, dim_factor=1, num_layer=3,
          time_decay_weight=shape.time_decay_weight, use_matrix=False, beginning_time=0.0, not_scale=False, enforce_dim=24)
chunked = WalkProjectionOracle(**kw)
seq = WalkProjectionOracle(**kw, p0=chunked.P[0])
exact = [p.astype(np.float64) for p in seq.P]
now = 0.0
print(f'# batch {B} edges, chunk {a.chunk}, d=24 (the error statistics do not depend on d)')
for k, (s, d, t) in enumerate(edge_stream(shape, B, a.batches, seed=1234)):
    chunked.update(s, d, t, giant_chunk=a.chunk)
    seq.update(s, d, t)
    w = edge_weights(t, t[-1], shape.time_decay_weight).astype(np.float64)
    c = decay_factors(shape.time_decay_weight, t[-1], now, 3).astype(np.float64)
    for i in range(1, 4):
        exact[i] = exact[i] * c[i - 1]
    for i in range(3, 0, -1):
        to_src, to_dst = exact[i - 1][d] * w[:, None], exact[i - 1][s] * w[:, None]
        np.add.at(exact[i], s, to_src)
        np.add.at(exact[i], d, to_dst)
    now = t[-1]
    top = int(np.bincount(np.concatenate([s, d])).max())
    for i in range(1, 4):
        scale = np.abs(exact[i]).max(axis=1, keepdims=True) + 1e-300
        e_seq = float((np.abs(seq.P[i] - exact[i]) / scale).max())
        e_chk = float((np.abs(chunked.P[i] - exact[i]) / scale).max())
        e_between = float((np.abs(seq.P[i] - chunked.P[i]) / scale).max())
        print(f'batch {k} (top hub {top} messages) layer {i}: reference order vs exact {e_seq:.2e} | chunked vs exact '
              f'{e_chk:.2e} | chunked vs reference order {e_between:.2e}', flush=True)
