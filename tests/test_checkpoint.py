"""Packed / sharded checkpoint format (tpnet_b200/checkpoint.py): host logic, runs on CPU."""
import numpy as np
import pytest
import torch

from tpnet_b200 import RandomProjectionModule
from tpnet_b200 import checkpoint as ck


def make(node_num=53, seed=0, **kw):
    torch.manual_seed(seed)
    args = dict(node_num=node_num, edge_num=500, dim_factor=3, num_layer=2, time_decay_weight=1e-6, device='cpu',
                use_matrix=False, beginning_time=np.float64(12.5), not_scale=False, enforce_dim=-1)
    args.update(kw)
    m = RandomProjectionModule(**args)
    with torch.no_grad():
        for i in (1, 2):
            m.random_projections[i].normal_()
        m.now_time.fill_(99.25)
    m._after_external_write()
    return m


def test_roundtrip_single_file(tmp_path):
    a, b = make(seed=1), make(seed=2)
    path = ck.save_checkpoint(a, str(tmp_path), 'ep3')
    assert path.endswith('ep3.rank0-of-1.pt')
    payload = torch.load(path, weights_only=True)
    assert payload['format'] == ck.FORMAT and payload['rows'] == 53 and 'state' not in payload
    import os
    assert os.path.getsize(ck.state_path(path)) == 53 * 3 * a.dim * 4   # stored once, raw fp32, pad columns dropped
    ck.load_checkpoint(b, str(tmp_path), 'ep3')
    for i in range(3):
        assert torch.equal(a.random_projections[i].data, b.random_projections[i].data)
    assert float(b.now_time) == 99.25 and b._now_host == 99.25 and float(b.begging_time) == 12.5
    assert all(torch.equal(p, q) for p, q in zip(a.mlp.state_dict().values(), b.mlp.state_dict().values()))
    assert b._is_packed() and not b._state[:, :, b.dim:].any()


@pytest.mark.parametrize('saved_world,load_world', [(1, 3), (3, 1), (2, 3), (4, 2)])
def test_resharding_by_node_id(tmp_path, saved_world, load_world):
    """Files written by G ranks load into any other number of ranks: node u <-> (rank u % G, row u // G)."""
    N, L, d = 41, 2, 6
    full = torch.randn(N, L + 1, d)
    head = {'w': torch.randn(3, 3)}
    for r in range(saved_world):
        torch.save(ck.pack_shard(full[r::saved_world].clone(), saved_world, r, N, 7.0, 1.0, head),
                   ck.shard_path(str(tmp_path), 't', r, saved_world))
    files = ck.list_shards(str(tmp_path), 't')
    assert len(files) == saved_world
    for r in range(load_world):
        rows = (N - r + load_world - 1) // load_world
        dst = torch.zeros(rows, L + 1, d + 2)                        # padded row stride, like the packed state
        got = sum(ck.scatter_shard(dst, load_world, r, torch.load(f, weights_only=True)) for f in files)
        assert got == rows
        assert torch.equal(dst[:, :, :d], full[r::load_world]) and not dst[:, :, d:].any()


def test_incomplete_or_mismatched_checkpoints_are_refused(tmp_path):
    N = 20
    full = torch.randn(N, 3, 4)
    torch.save(ck.pack_shard(full[0::2].clone(), 2, 0, N, 0.0, 0.0, {}), ck.shard_path(str(tmp_path), 'x', 0, 2))
    with pytest.raises(FileNotFoundError, match='ranks \\[1\\]'):
        ck.list_shards(str(tmp_path), 'x')
    with pytest.raises(FileNotFoundError):
        ck.list_shards(str(tmp_path), 'nothing')
    with pytest.raises(ValueError, match='owns'):
        ck.pack_shard(full[0::2][:-1].clone(), 2, 0, N, 0.0, 0.0, {})
    a = make()
    ck.save_checkpoint(a, str(tmp_path), 'y')
    with pytest.raises(ValueError, match='checkpoint is for'):
        ck.load_checkpoint(make(node_num=54), str(tmp_path), 'y')


def test_chunked_state_file_and_v1_compatibility(tmp_path):
    """Tiny chunk size: many chunks, ragged last one; and a v1 file (state inside the header) still loads."""
    a, b, c = make(seed=3, node_num=101), make(seed=4, node_num=101), make(seed=5, node_num=101)
    ck.save_checkpoint(a, str(tmp_path), 'small_chunks', chunk_bytes=7 * 3 * a.dim * 4)       # 7 rows per chunk
    ck.load_checkpoint(b, str(tmp_path), 'small_chunks', chunk_bytes=5 * 3 * a.dim * 4)
    for i in range(3):
        assert torch.equal(a.random_projections[i].data, b.random_projections[i].data)
    v1 = ck.pack_shard(a._state[:, :, :a.dim].clone(), 1, 0, 101, 99.25, 12.5, a.mlp.state_dict())
    torch.save(v1, ck.shard_path(str(tmp_path), 'old', 0, 1))
    ck.load_checkpoint(c, str(tmp_path), 'old')
    for i in range(3):
        assert torch.equal(a.random_projections[i].data, c.random_projections[i].data)
    # a truncated state file is refused
    with open(ck.state_path(ck.shard_path(str(tmp_path), 'small_chunks', 0, 1)), 'r+b') as fh:
        fh.truncate(100)
    with pytest.raises(ValueError, match='bytes, expected'):
        ck.load_checkpoint(b, str(tmp_path), 'small_chunks')
