#!/usr/bin/env python
"""Times tpn_head_forward per 100,000 pairs: packed-FFMA kernel (TPN_DEBUG_HEAD_FFMA) vs the tcgen05 kernel (default), and
torch's own fp32 nn.Sequential (cuBLAS) for reference.  CUDA events, 50 launches each after warm-up."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpnet_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = 'cuda:0'
torch.manual_seed(0)
mlp = torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.ReLU(), torch.nn.Linear(256, 64)).to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
x = torch.rand(n, 64, device=dev) * 12.0
y = torch.empty_like(x)
l1, l2 = mlp[0], mlp[2]
stream = torch.cuda.current_stream().cuda_stream


def run():
    rc = lib.tpn_head_forward(x.data_ptr(), n, None, 64, 256, l1.weight.data_ptr(), l1.bias.data_ptr(),
                              l2.weight.data_ptr(), l2.bias.data_ptr(), y.data_ptr(), stream)
    assert rc == 0, rc


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


out = {'pairs': n, 'flop': 2 * n * (64 * 256 + 256 * 64)}
with torch.no_grad():
    ref64 = mlp.double()(x.double())
    mlp.float()
    for name, flag in (('ffma', 8), ('tensor', 0)):
        old = lib.tpn_set_debug_flags(flag)
        run()
        torch.cuda.synchronize()
        err = float((y.double() - ref64).abs().max())
        us = timed(run)
        lib.tpn_set_debug_flags(old)
        out[name] = {'us': us, 'TFLOPs': out['flop'] / us / 1e6, 'max_abs_err_vs_f64': err}
    t32 = mlp(x)
    out['torch_fp32'] = {'us': timed(lambda: mlp(x)), 'max_abs_err_vs_f64': float((t32.double() - ref64).abs().max())}
print(json.dumps(out))
