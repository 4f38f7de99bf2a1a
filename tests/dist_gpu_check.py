"""Multi-GPU check of the sharded path, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_gpu_check.py

Every rank runs ShardedRandomProjection on the same replicated stream; rank 0 also runs the
single-GPU module.  The sharded state must equal the single-GPU state BIT FOR BIT (same
kernels, same per-row accumulation order); pair-wise features within the fp32 dot-product
tolerance (different launch shapes pick different reduction trees).  Both data planes: 'peer' (device-side
routing + NVLink pulls + flag barriers; bit-exact in lazy mode too) and 'nccl' (host plan + all_to_all).
Any world size (2, 4, 8): the driver of the check is scripts/gpu_multi*.sh / tests/test_gpu_sharded.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tpnet_b200 import RandomProjectionModule  # noqa: E402
from tpnet_b200.sharded import ShardedRandomProjection  # noqa: E402


def stream(rng, N, B, nb, skew):
    t = 0.0
    for _ in range(nb):
        s = 1 + (rng.zipf(skew, B) - 1) % (N - 1)
        d = 1 + (rng.zipf(skew, B) - 1) % (N - 1)
        ts = np.sort(t + rng.random(B) * 500.0)
        t = ts[-1]
        yield s.astype(np.int64), d.astype(np.int64), ts


def main():
    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    ok = True
    cases = [(exchange, mode, 'reference') for exchange in ('peer', 'nccl') for mode in ('eager', 'lazy')]
    cases.append(('peer', 'lazy', 'chunked'))
    for exchange, mode, acc in cases:
        for (N, B, dim, L) in [(403, 200, 20, 3), (5003, 40000, 24, 2), (997, 3000, 140, 3), (301, 9000, 64, 3)]:
            kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=1e-4,
                      device=str(dev), use_matrix=False, beginning_time=np.float64(0.0), not_scale=False,
                      enforce_dim=dim)
            torch.manual_seed(7)
            sh = ShardedRandomProjection(decay_mode=mode, ext_rows=2 * B + 16, exchange=exchange, state_device=dev,
                                         accumulation=acc, giant_chunk=512, **kw).to(dev)
            ref = None
            if rank == 0:
                torch.manual_seed(7)
                ref = RandomProjectionModule(decay_mode=mode, accumulation=acc, giant_chunk=512, **kw).to(dev)
            rng = np.random.default_rng(N)
            for s, d, t in stream(rng, N, B, 4, 1.3):
                sh.update(s, d, t)
                if ref is not None:
                    ref.update(s, d, t)
            full = sh.gather_global()
            n = 3001
            a = rng.integers(0, N, n).astype(np.int64)
            b = rng.integers(0, N, n).astype(np.int64)
            keep, feat = sh.pair_wise_gram(a, b)
            if rank == 0:
                ref.materialize()
                for i in range(L + 1):
                    want_i = ref.random_projections[i].data
                    if mode == 'eager' or exchange == 'peer':
                        # peer data plane: cached remote rows carry their owner's stamps -> bit-exact in lazy mode too
                        same = torch.equal(full[i], want_i)
                    else:       # nccl + lazy: a received row is rescaled by the sender and again by the reader
                        # (two roundings), the single-GPU run rescales once: rtol 1e-5
                        same = torch.allclose(full[i], want_i, rtol=1e-5, atol=1e-6 * max(float(want_i.abs().max()), 1.0))
                    ok &= same
                    if not same:
                        err = (full[i] - ref.random_projections[i].data).abs().max().item()
                        print(f'MISMATCH exchange={exchange} mode={mode} acc={acc} N={N} B={B} layer {i}: max abs err {err}')
                keep_t = keep if isinstance(keep, torch.Tensor) else torch.from_numpy(keep).to(dev)
                want = ref.pair_wise_gram(a, b)[keep_t]
                close = torch.allclose(feat, want, rtol=1e-5, atol=2e-5)
                ok &= close
                if not close:
                    print(f'PAIRWISE MISMATCH mode={mode} N={N}: {(feat - want).abs().max().item()}')
            sh.check_errors()
            sh.close()                               # unmap the peers' buffers before anyone frees them
            del sh, full
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.broadcast(flag, 0)
            ok = bool(flag.item())
            if rank == 0:
                print(f'exchange={exchange} mode={mode} acc={acc} N={N} B={B} d={dim} L={L} world={world}: '
                      f'{"ok" if ok else "FAILED"}')
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)
    if rank == 0:
        print('sharded == single GPU: PASS')


if __name__ == '__main__':
    main()
