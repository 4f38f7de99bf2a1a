#!/usr/bin/env python
"""Timeline of the first giant work item of the hub walker.  Needs a profiling build:
TPN_EXTRA_NVCC_FLAGS=-DTPN_HUB2_TIMELINE python -m tpnet_b200.build --force"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpnet_b200 import RandomProjectionModule, _lib  # noqa: E402

dev = 'cuda:0'
N, B = 1_000_001, int(os.environ.get('PROBE_HUB', 100_000))
m = RandomProjectionModule(node_num=N, edge_num=10**9, dim_factor=10, num_layer=3, time_decay_weight=1e-7, device=dev,
                           use_matrix=False, beginning_time=np.float64(0.0), not_scale=False, enforce_dim=-1,
                           decay_mode='lazy', init_p0=False, state_device=dev).to(dev)
m.random_projections[0].data.normal_(0, 0.07)
rng = np.random.default_rng(0)
others = rng.permutation(np.arange(2, N))[:B].astype(np.int64)
src = np.full(B, 1, dtype=np.int64)
t = 0.0
for _ in range(3):
    ts = np.sort(t + rng.random(B) * 3000.0)
    t = ts[-1]
    m.update(src, others, ts)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
n = 8 * 256 * 8
buf = (ctypes.c_ulonglong * n)()
lib.tpn_debug_hub_timeline.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
assert lib.tpn_debug_hub_timeline(buf, n) == 0
a = np.array(buf[:], dtype=np.int64).reshape(8, 256, 8)
t0 = a[0, 0, 0]
print('consumer (warp 0) per stage: [before wait, after wait, after chain, after release]')
for b in range(200, 206):
    print('  stage', b, (a[0, b, :4] - t0).tolist())
d = a[0, 40:250]
print('consumer, stages 40-120 :', 'period', np.diff(a[0, 40:120, 0]).mean(), 'wait', (a[0, 40:120, 1] - a[0, 40:120, 0]).mean())
print('consumer, stages 170-250:', 'period', np.diff(a[0, 170:250, 0]).mean(), 'wait', (a[0, 170:250, 1] - a[0, 170:250, 0]).mean())
print('consumer: period', np.diff(d[:, 0]).mean(), 'wait', (d[:, 1] - d[:, 0]).mean(), 'chain', (d[:, 2] - d[:, 1]).mean(),
      'release', (d[:, 3] - d[:, 2]).mean())
print('producer warp 1 per pass: [start, meta issued, rows issued, after wait-empty, rows arrived, stored, published]')
for p in range(8, 14):
    print('  pass', p, (a[1, p, :7] - t0).tolist())
for w in range(1, 8):
    x = a[w, 6:28]
    print(f'producer w{w}: period', np.diff(x[:, 0]).mean(), '| meta issue', (x[:, 1] - x[:, 0]).mean(), '| shuffles+row issue',
          (x[:, 2] - x[:, 1]).mean(), '| wait empty', (x[:, 3] - x[:, 2]).mean(), '| rows arrive', (x[:, 4] - x[:, 3]).mean(),
          '| scale+store', (x[:, 5] - x[:, 4]).mean(), '| publish', (x[:, 6] - x[:, 5]).mean())
