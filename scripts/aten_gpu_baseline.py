#!/usr/bin/env python
"""Comparison point, not part of bench.py: the reference's own op sequence (stock ATen kernels — index, mul,
scatter_add_ with float atomics, batched GEMM, log, Linear head; oracle/cpu_port.py with device='cuda:0') on the
same GPU, i.e. what the unmodified reference does with `--gpu 0`.  Power-law workload on a 10x smaller node set
(the reference's eager decay rewrites the whole state every update), numpy inputs, wall clock.

    python scripts/aten_gpu_baseline.py [--batch 100000] [--steps 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=bench.PL_BATCH)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--scale-down', type=int, default=10)
a = ap.parse_args()
shape = bench.SHAPES['powerlaw']
out = {'workload': f'power-law d={shape.dim} L={shape.num_layer} batch {a.batch}, {a.scale_down}x fewer nodes'}
for dev in ('cuda:0', 'cpu'):
    sec, n = bench.cpu_port_powerlaw(shape, a.batch, 2, a.steps, os.cpu_count() or 1, a.scale_down, device=dev)
    out[dev] = {'edges_per_s': a.batch / sec, 'ms_per_step': sec * 1e3, 'nodes': n}
    if dev != 'cpu':
        torch.cuda.empty_cache()
print(json.dumps(out))
