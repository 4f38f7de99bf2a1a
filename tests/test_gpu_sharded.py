"""GPU suite: the sharded path on >= 2 GPUs (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_single_rank_equals_plain_module():
    """world_size 1 (no process group): the message-mode update + extension-row plumbing must
    reproduce the plain module bit for bit."""
    import numpy as np
    from tpnet_b200 import RandomProjectionModule
    from tpnet_b200.sharded import ShardedRandomProjection
    dev = 'cuda:0'
    for mode in ('eager', 'lazy'):
        kw = dict(node_num=301, edge_num=3000, dim_factor=1, num_layer=3, time_decay_weight=1e-4, device=dev,
                  use_matrix=False, beginning_time=np.float64(0.0), not_scale=False, enforce_dim=28)
        torch.manual_seed(3)
        sh = ShardedRandomProjection(decay_mode=mode, ext_rows=64, **kw).to(dev)
        torch.manual_seed(3)
        ref = RandomProjectionModule(decay_mode=mode, **kw).to(dev)
        rng = np.random.default_rng(1)
        t = 0.0
        for B in (200, 3000, 40000):
            s = rng.integers(1, 301, B).astype(np.int64)
            d = rng.integers(1, 301, B).astype(np.int64)
            ts = np.sort(t + rng.random(B) * 100.0)
            t = ts[-1]
            sh.update(s, d, ts, plan=sh.plan_update(s, d, ts))      # explicit plan: the message-mode path
            ref.update(s, d, ts)
        sh.materialize(); ref.materialize()
        for i in range(4):
            assert torch.equal(sh.random_projections[i].data[:301], ref.random_projections[i].data), (mode, i)
        a = rng.integers(0, 301, 500).astype(np.int64)
        b = rng.integers(0, 301, 500).astype(np.int64)
        keep, feat = sh.pair_wise_gram(a, b, plan=sh.plan_pairs(a, b))
        assert len(keep) == 500 and torch.equal(feat, ref.pair_wise_gram(a, b))
        keep, feat = sh.pair_wise_gram(a, b)                        # world 1 without a plan: the plain path
        assert len(keep) == 500 and torch.equal(feat, ref.pair_wise_gram(a, b))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_sharded_equals_single_gpu_two_ranks():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tests', 'dist_gpu_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'PASS' in r.stdout
