"""GPU suite of the `recent` neighbour sampler (tpn_sampler_recent) against the oracle
(oracle/neighbor_sampler.py) and the reference fixture.  Integer / index work: everything is bit-exact."""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR
from oracle.neighbor_sampler import RecentNeighborOracle
from tpnet_b200 import RandomProjectionModule
from tpnet_b200.neighbor_sampler import RecentNeighborSampler, get_neighbor_sampler

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_sampler_matches_reference_fixture():
    z = np.load(os.path.join(GOLDEN_DIR, 'sampler_tiny.npz'))
    s = RecentNeighborSampler(z['src'], z['dst'], z['eid'], z['t'], DEV, num_nodes=int(z['num_nodes']))
    for k in range(3):
        got = s.get_historical_neighbors(z[f'q{k}_nodes'], z[f'q{k}_times'], int(z[f'q{k}_K']))
        for x, name in zip(got, ('nbr', 'eid', 't')):
            ref = z[f'q{k}_{name}']
            assert x.shape == ref.shape and x.dtype == ref.dtype and np.array_equal(x, ref), (k, name)


@pytest.mark.parametrize('K', [1, 20, 33])
def test_sampler_vs_oracle_large(K):
    rng = np.random.default_rng(K)
    N, E, n = 3000, 60000, 7001
    src = 1 + (rng.zipf(1.3, E) - 1) % (N - 1)
    dst = 1 + (rng.zipf(1.3, E) - 1) % (N - 1)
    t = np.sort(np.floor(rng.random(E) * 5000.0) / 4.0)                 # many equal timestamps
    eid = np.arange(1, E + 1)
    o = RecentNeighborOracle(src, dst, eid, t, N)

    class Data:                                                          # the fields of the reference's Data object
        src_node_ids, dst_node_ids, edge_ids, node_interact_times = src, dst, eid, t
    s = get_neighbor_sampler(Data, 'recent', device=DEV)
    assert s.num_nodes == int(max(src.max(), dst.max())) + 1
    qn = rng.integers(0, s.num_nodes, n).astype(np.int64)
    qt = np.where(rng.random(n) < 0.5, np.floor(rng.random(n) * 5200.0) / 4.0, rng.random(n) * 1300.0)
    qt[:10] = -1.0                                                       # before everything
    qt[10:20] = 1e9                                                      # after everything
    ref = o.get_historical_neighbors(qn, qt, K)
    got = s.get_historical_neighbors(qn, qt, K)
    for x, y in zip(got, ref):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    assert not got[0][:10].any() and not got[0][qn == 0].any()
    # device-resident queries and results
    dn, de, dt = s.get_historical_neighbors(torch.from_numpy(qn).to(DEV), torch.from_numpy(qt).to(DEV), K, as_tensors=True)
    assert dn.is_cuda and dn.dtype == torch.int64 and dt.dtype == torch.float64
    assert np.array_equal(dn.cpu().numpy(), ref[0]) and np.array_equal(de.cpu().numpy(), ref[1])
    assert np.array_equal(dt.cpu().numpy(), ref[2])
    # nothing to do / ids outside the graph have no history
    e = s.get_historical_neighbors(qn[:0], qt[:0], K)
    assert e[0].shape == (0, K)
    # host ids outside the graph: IndexError, as the reference's per-node list lookup (utils/utils.py:177) raises;
    # device-resident ids cannot raise from a kernel: such queries have no history (all-zero rows)
    with pytest.raises(IndexError):
        s.get_historical_neighbors(np.array([s.num_nodes + 5, 1], dtype=np.int64), np.array([1e9, 1e9]), K)
    far = s.get_historical_neighbors(torch.tensor([s.num_nodes + 5, -3], dtype=torch.int64, device=DEV),
                                     torch.tensor([1e9, 1e9], dtype=torch.float64, device=DEV), K)
    assert not far[0].any() and not far[1].any() and not far[2].any()
    with pytest.raises(NotImplementedError):
        get_neighbor_sampler(Data, 'uniform', device=DEV)


def test_sampled_ids_feed_the_structured_encoder_call_on_device():
    """Sampler output (device tensors) -> get_neighbor_pair_wise_feature, no host round trip: same features as
    the numpy path."""
    rng = np.random.default_rng(4)
    N, E, B, K = 400, 6000, 64, 20
    src = rng.integers(1, N, E)
    dst = rng.integers(1, N, E)
    t = np.sort(rng.random(E) * 1000.0)
    s = RecentNeighborSampler(src, dst, np.arange(1, E + 1), t, DEV, num_nodes=N)
    torch.manual_seed(0)
    m = RandomProjectionModule(node_num=N, edge_num=E + 1, dim_factor=10, num_layer=3, time_decay_weight=1e-6, device=DEV,
                               use_matrix=False, beginning_time=np.float64(0.0), not_scale=False, enforce_dim=-1).to(DEV)
    for i in range(0, 3000, 500):
        m.update(src[i:i + 500].astype(np.int64), dst[i:i + 500].astype(np.int64), t[i:i + 500])
    bs, bd = src[3000:3000 + B].astype(np.int64), dst[3000:3000 + B].astype(np.int64)
    rows = np.concatenate([bs, bd])
    times = np.tile(t[3000:3000 + B], 2)
    nbr_np, _, _ = s.get_historical_neighbors(rows, times, K)
    nbr_dev, _, _ = s.get_historical_neighbors(rows, times, K, as_tensors=True)
    s2, d2 = np.tile(bs, 2), np.tile(bd, 2)
    with torch.no_grad():
        a = m.get_neighbor_pair_wise_feature(nbr_np, s2, d2)
        b = m.get_neighbor_pair_wise_feature(nbr_dev, torch.from_numpy(s2).to(DEV), torch.from_numpy(d2).to(DEV))
    assert a.shape == (2 * B, K, 2 * m.pair_wise_feature_dim) and torch.equal(a, b)
    m.check_errors()
