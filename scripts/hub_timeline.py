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
print('consumer (warp 0), every 8th stage of the first chain item: [before wait, after wait, after chain, after release]')
nst = min(256, (B + 127) // 128 // 8)
for lo in range(0, nst, 32):
    d = a[0, lo:min(lo + 32, nst)]
    if len(d) < 2:
        break
    print('consumer, stages %4d-%4d: period per stage %.0f  wait %.0f  chain %.0f  release %.0f' % (
        8 * lo, 8 * (lo + len(d)), np.diff(d[:, 0]).mean() / 8, (d[:, 1] - d[:, 0]).mean(), (d[:, 2] - d[:, 1]).mean(),
        (d[:, 3] - d[:, 2]).mean()))
