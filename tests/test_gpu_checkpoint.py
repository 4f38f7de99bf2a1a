"""GPU suite: packed / sharded checkpoints (tpnet_b200/checkpoint.py) with a GPU-resident state:
chunked save of a lazily-decayed state in mid-stream, reload into a fresh module, and re-sharding into two ranks
(peer data plane, both ranks in this process) that then continue the stream bit-identically."""
import numpy as np
import pytest
import torch

from tpnet_b200 import RandomProjectionModule
from tpnet_b200 import checkpoint as ck

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _stream(rng, N, B, nb, t0=0.0):
    t = t0
    for _ in range(nb):
        s = (1 + (rng.zipf(1.3, B) - 1) % (N - 1)).astype(np.int64)
        d = (1 + (rng.zipf(1.3, B) - 1) % (N - 1)).astype(np.int64)
        ts = np.sort(t + rng.random(B) * 3000.0)
        t = ts[-1]
        yield s, d, ts


@pytest.mark.parametrize('mode', ['eager', 'lazy'])
def test_gpu_resident_save_reload_and_reshard(tmp_path, mode):
    from test_gpu_sharded import _each, _global_layers, _sim_ranks, _update_all
    N, L = 4003, 3
    kw = dict(node_num=N, edge_num=40000, dim_factor=10, num_layer=L, time_decay_weight=1e-5, device=DEV,
              use_matrix=False, beginning_time=np.float64(0.0), not_scale=False, enforce_dim=-1)
    torch.manual_seed(2)
    a = RandomProjectionModule(decay_mode=mode, **kw).to(DEV)
    rng = np.random.default_rng(8)
    first = list(_stream(rng, N, 3000, 4))
    for s, d, t in first:
        a.update(s, d, t)
    row_bytes = (L + 1) * a.dim * 4
    path = ck.save_checkpoint(a, str(tmp_path), 'mid', chunk_bytes=257 * row_bytes)        # 16 chunks, ragged tail
    assert ck.state_path(path).endswith('mid.rank0-of-1.state')
    # (1) a fresh module continues the stream exactly like the original
    torch.manual_seed(3)
    b = RandomProjectionModule(decay_mode=mode, **kw).to(DEV)
    ck.load_checkpoint(b, str(tmp_path), 'mid', chunk_bytes=100 * row_bytes)
    assert float(b.now_time) == float(a.now_time) == first[-1][2][-1]
    # (2) the same checkpoint re-sharded into two ranks
    ranks, streams = _sim_ranks(2, kw, mode, ext_rows=2 * N)
    for m in ranks:
        ck.load_checkpoint(m, str(tmp_path), 'mid')
        m._state_written()                                   # rows changed under the remote-row caches
    rest = list(_stream(rng, N, 3000, 3, t0=first[-1][2][-1]))
    for s, d, t in rest:
        a.update(s, d, t)
        b.update(s, d, t)
        _update_all(ranks, streams, s, d, t)
    a.materialize(); b.materialize()
    full = _global_layers(ranks, N, L, streams)
    for i in range(L + 1):
        assert torch.equal(a.random_projections[i].data, b.random_projections[i].data), (mode, i)
        if mode == 'eager':
            assert torch.equal(full[i], a.random_projections[i].data), ('resharded', mode, i)
        else:       # the reloaded ranks restart their decay log at the checkpoint: one extra rounding per row
            scale = max(float(a.random_projections[i].data.abs().max()), 1.0)
            assert torch.allclose(full[i], a.random_projections[i].data, rtol=1e-5, atol=1e-6 * scale), ('resharded', mode, i)
    # (3) two ranks save, one module loads (8 -> 1 style re-assembly)
    for m, st in zip(ranks, streams):
        with torch.cuda.stream(st):
            ck.save_checkpoint(m, str(tmp_path), 'two')
    torch.cuda.synchronize()
    torch.manual_seed(4)
    c = RandomProjectionModule(decay_mode=mode, **kw).to(DEV)
    ck.load_checkpoint(c, str(tmp_path), 'two')
    for i in range(L + 1):
        assert torch.equal(c.random_projections[i].data, full[i]), ('reassembled', mode, i)
