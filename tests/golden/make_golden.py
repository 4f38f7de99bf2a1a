#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference class
(``/root/reference/models/TPNet.py:9-157``) on CPU, and pins the oracle to it.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

For every case it
  1. drives the imported reference ``RandomProjectionModule`` (device='cpu');
  2. drives ``oracle.walk_projection.WalkProjectionOracle`` with the reference's
     own torch-computed edge weights and asserts BIT equality of every layer
     after every batch (pins decay / ordering / layer semantics);
  3. drives the oracle with its own weights and asserts closeness (<= 4 ulp of
     accumulated drift is all torch's 1-ulp exp can cause);
  4. drives ``oracle.cpu_port.CpuWalkProjection`` and asserts BIT equality;
  5. stores inputs + reference outputs (+ the reference's weights) as fixtures.
The fixtures are what travels: nothing at test time reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

from models.TPNet import RandomProjectionModule as RefModule  # noqa: E402  (the reference)

from oracle.cpu_port import CpuWalkProjection  # noqa: E402
from oracle.walk_bruteforce import random_temporal_graph, sum_walk_matrices  # noqa: E402
from oracle.walk_projection import WalkProjectionOracle  # noqa: E402


def ref_weights(times, lam):
    # same expression as TPNet.py:76-78, evaluated by torch on the CPU
    t_last = times[-1]
    tf = torch.from_numpy(times).to(dtype=torch.float)
    return torch.exp(-lam * (t_last - tf)).numpy().copy()


def layers_of(ref):
    return [p.data.numpy().copy() for p in ref.random_projections]


def run_case(name, *, node_num, edge_num, dim_factor, num_layer, lam, use_matrix, not_scale, enforce_dim,
             batches, pair_a, pair_b, beginning_time, seed, backup_after=None):
    torch.manual_seed(seed)
    kw = dict(node_num=node_num, edge_num=edge_num, dim_factor=dim_factor, num_layer=num_layer,
              time_decay_weight=lam, use_matrix=use_matrix, beginning_time=np.float64(beginning_time),
              not_scale=not_scale, enforce_dim=enforce_dim)
    ref = RefModule(device='cpu', **kw)
    ref.mlp = torch.nn.Identity()                       # fixtures hold the INPUT of the trainable head
    p0 = ref.random_projections[0].data.numpy().copy()
    pinned = WalkProjectionOracle(p0=None if use_matrix else p0, **kw)      # fed the reference's weights
    free = WalkProjectionOracle(p0=None if use_matrix else p0, **kw)        # its own weights
    port = None
    if not use_matrix:
        port = CpuWalkProjection(node_num, ref.dim, num_layer, lam, beginning_time, not_scale, with_mlp=False)
        port.layers[0] = torch.from_numpy(p0.copy())

    out = {'p0': p0, 'n_batches': np.int64(len(batches)), 'dim': np.int64(ref.dim)}
    for k, v in kw.items():
        out['cfg_' + k] = np.asarray(v)
    saved = None
    for b, (s, d, t) in enumerate(batches):
        w = ref_weights(t, lam)
        ref.update(s, d, t)
        pinned.update(s, d, t, weights=w)
        free.update(s, d, t)
        got = layers_of(ref)
        for i in range(num_layer + 1):
            assert np.array_equal(got[i], pinned.P[i]), f'{name}: oracle != reference, batch {b} layer {i}'
            np.testing.assert_allclose(free.P[i], got[i], rtol=2e-6, atol=1e-7)
        assert np.float64(ref.now_time.item()) == pinned.now_time
        if port is not None:
            port.update(s, d, t)
            for i in range(num_layer + 1):
                assert np.array_equal(got[i], port.layers[i].numpy()), f'{name}: port != reference'
        out[f'b{b}_src'], out[f'b{b}_dst'], out[f'b{b}_t'], out[f'b{b}_w'] = s, d, t, w
        if b == 0:
            for i in range(1, num_layer + 1):
                out[f'after0_P{i}'] = got[i]
        if backup_after is not None and b == backup_after:
            saved = ref.backup_random_projections()
            saved_or = pinned.backup()
            out['backup_now'] = np.float64(saved[0].item())
            for i in range(num_layer):
                out[f'backup_P{i + 1}'] = saved[1][i].numpy().copy()
    got = layers_of(ref)
    for i in range(1, num_layer + 1):
        out[f'final_P{i}'] = got[i]
    out['final_now'] = np.float64(ref.now_time.item())

    with torch.no_grad():
        feat = ref.get_pair_wise_feature(pair_a, pair_b).numpy().copy()
    out['pair_a'], out['pair_b'], out['pair_feat'] = pair_a, pair_b, feat
    o_feat = pinned.pair_wise_gram(pair_a, pair_b)
    scale = pinned.pair_norm_bound(pair_a, pair_b)
    if not_scale:
        assert np.all(np.abs(o_feat - feat) <= 1e-5 * np.abs(feat) + 2e-6 * scale), f'{name}: pair-wise gram'
    else:
        np.testing.assert_allclose(o_feat, feat, rtol=1e-5, atol=2e-6)
    if port is not None:
        with torch.no_grad():
            assert np.array_equal(port.gram_features(pair_a, pair_b).numpy(), feat), f'{name}: port pair-wise'
    gathered = ref.get_random_projections(pair_a)
    for i in range(num_layer + 1):
        assert np.array_equal(gathered[i].numpy(), pinned.get_random_projections(pair_a)[i])

    # the encoder's structured call, TPNet.py:313-324 verbatim (m rows, K neighbours, front-padded with id 0)
    nrng = np.random.default_rng(1000 + seed)
    m_rows, K = 10, 5 if num_layer != 3 else 6
    nbr = nrng.integers(1, node_num, (m_rows, K)).astype(np.int64)
    nbr[0, :] = 0
    nbr[1, :3] = 0
    nsrc = nrng.integers(1, node_num, m_rows).astype(np.int64)
    ndst = nrng.integers(1, node_num, m_rows).astype(np.int64)
    nbr[2, -1] = nsrc[2]                                   # a neighbour that is the row's own source
    ndst[3] = nsrc[3]                                      # src == dst
    with torch.no_grad():
        concat_neighbor_random_features = ref.get_pair_wise_feature(
            src_node_ids=np.tile(nbr.reshape(-1), 2),
            dst_node_ids=np.concatenate([np.repeat(nsrc, K), np.repeat(ndst, K)]))
        neighbor_random_features = torch.cat(
            [concat_neighbor_random_features[:m_rows * K], concat_neighbor_random_features[m_rows * K:]],
            dim=1).reshape(m_rows, K, -1)
    nfeat = neighbor_random_features.numpy().copy()
    out['nbr'], out['nbr_src'], out['nbr_dst'], out['nbr_feat'] = nbr, nsrc, ndst, nfeat
    o_nfeat = pinned.neighbor_pair_wise_gram(nbr, nsrc, ndst)
    assert o_nfeat.shape == nfeat.shape
    la, lb = pinned.neighbor_pair_lists(nbr, nsrc, ndst)
    nscale = pinned.pair_norm_bound(la, lb)
    nscale = np.concatenate([nscale[:m_rows * K], nscale[m_rows * K:]], axis=1).reshape(m_rows, K, -1)
    if not_scale:
        assert np.all(np.abs(o_nfeat - nfeat) <= 1e-5 * np.abs(nfeat) + 2e-6 * nscale), f'{name}: neighbour gram'
    else:
        np.testing.assert_allclose(o_nfeat, nfeat, rtol=1e-5, atol=2e-6)

    if saved is not None:
        ref.reload_random_projections(saved)
        pinned.reload(saved_or)
        assert np.float64(ref.now_time.item()) == pinned.now_time
        # after a reload the stream continues from the backed-up state: replay the last batch
        s, d, t = batches[-1]
        # time goes "backwards" relative to the pre-reload clock; relative to the reloaded clock it is forward
        w = ref_weights(t, lam)
        ref.update(s, d, t)
        pinned.update(s, d, t, weights=w)
        got = layers_of(ref)
        for i in range(1, num_layer + 1):
            assert np.array_equal(got[i], pinned.P[i])
            out[f'after_reload_P{i}'] = got[i]
        out['after_reload_w'] = w
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: ok  dim={ref.dim}  {os.path.getsize(path) / 1024:.1f} KiB')
    return ref, pinned


def bipartite_batches(rng, n_src, n_dst, n_batches, B, t0, span, skew):
    out, t = [], t0
    for _ in range(n_batches):
        s = 1 + (rng.zipf(skew, B) - 1) % n_src
        d = 1 + n_src + (rng.zipf(skew, B) - 1) % n_dst
        ts = np.sort(t + rng.random(B) * span)
        t = ts[-1]
        out.append((s.astype(np.int64), d.astype(np.int64), ts.astype(np.float64)))
    return out


def main():
    rng = np.random.default_rng(20241017)

    # 1. Wikipedia-like: bipartite, L=2, ragged d (14), float timestamps, pad id 0 in the pair list
    n_src, n_dst = 45, 14
    N = n_src + n_dst + 1
    batches = bipartite_batches(rng, n_src, n_dst, 12, 20, 1000.0, 900.0, 1.3)
    a = rng.integers(0, N, 64).astype(np.int64); b = rng.integers(0, N, 64).astype(np.int64)
    a[:4] = 0; b[2:6] = 0
    run_case('wiki_tiny', node_num=N, edge_num=600, dim_factor=2, num_layer=2, lam=1e-3, use_matrix=False,
             not_scale=False, enforce_dim=-1, batches=batches, pair_a=a, pair_b=b, beginning_time=batches[0][2][0],
             seed=1, backup_after=7)

    # 2. Flights-like: non-bipartite, L=3, day-quantised timestamps (many equal), heavy duplicates,
    #    self loops and repeated edges inside one batch, raw (not_scale) features
    N = 40
    day = 0
    batches = []
    for k in range(10):
        B = 24
        s = 1 + (rng.zipf(1.5, B) - 1) % (N - 1)
        d = 1 + (rng.zipf(1.5, B) - 1) % (N - 1)
        s[3] = d[3]                      # self loop
        s[5], d[5] = s[4], d[4]          # repeated edge
        s[7], d[7] = d[6], s[6]          # reversed repeat
        if k % 3 == 2:
            day += 1
        ts = np.full(B, day * 86400.0)
        if k % 3 == 1:
            ts[B // 2:] += 86400.0       # a batch that straddles two days
        batches.append((s.astype(np.int64), d.astype(np.int64), ts))
    a = rng.integers(1, N, 48).astype(np.int64); b = rng.integers(1, N, 48).astype(np.int64)
    run_case('flights_tiny', node_num=N, edge_num=500, dim_factor=10, num_layer=3, lam=1e-6, use_matrix=False,
             not_scale=True, enforce_dim=10, batches=batches, pair_a=a, pair_b=b, beginning_time=0.0, seed=2,
             backup_after=4)

    # 3. Reddit-like: L=3 default, d multiple of 4, scaled features, strong decay
    n_src, n_dst = 70, 9
    N = n_src + n_dst + 1
    batches = bipartite_batches(rng, n_src, n_dst, 16, 32, 0.0, 4000.0, 1.2)
    a = rng.integers(0, N, 96).astype(np.int64); b = rng.integers(0, N, 96).astype(np.int64)
    run_case('reddit_tiny', node_num=N, edge_num=2000, dim_factor=4, num_layer=3, lam=1e-4, use_matrix=False,
             not_scale=False, enforce_dim=-1, batches=batches, pair_a=a, pair_b=b, beginning_time=0.0, seed=3)

    # 4. Single-layer and 4-layer corner cases
    batches = bipartite_batches(rng, 20, 6, 5, 8, 10.0, 50.0, 1.4)
    a = rng.integers(0, 27, 16).astype(np.int64); b = rng.integers(0, 27, 16).astype(np.int64)
    run_case('one_layer', node_num=27, edge_num=100, dim_factor=3, num_layer=1, lam=1e-2, use_matrix=False,
             not_scale=False, enforce_dim=-1, batches=batches, pair_a=a, pair_b=b, beginning_time=10.0, seed=4)
    run_case('four_layer', node_num=27, edge_num=100, dim_factor=3, num_layer=4, lam=1e-2, use_matrix=False,
             not_scale=False, enforce_dim=9, batches=batches, pair_a=a, pair_b=b, beginning_time=10.0, seed=5)

    # 5. Known-answer test: explicit walk matrices (use_matrix) vs exhaustive walk enumeration,
    #    under the notebook's batch preconditions (cell 5): integer batch timestamps
    N, E, L, lam, B = 24, 120, 3, 1e-4, 6
    s, d, _ = random_temporal_graph(N, E, rng)
    t = np.repeat(np.arange(1, E // B + 1), B).astype(np.float64)
    batches = [(s[i:i + B], d[i:i + B], t[i:i + B]) for i in range(0, E, B)]
    a = rng.integers(0, N, 16).astype(np.int64); b = rng.integers(0, N, 16).astype(np.int64)
    ref, orc = run_case('matrix_kat', node_num=N, edge_num=E, dim_factor=1, num_layer=L, lam=lam, use_matrix=True,
                        not_scale=True, enforce_dim=-1, batches=batches, pair_a=a, pair_b=b, beginning_time=0.0,
                        seed=6)
    brute = sum_walk_matrices(s, d, t, L, lam, N)
    for j in range(L + 1):
        # notebook cells 4/6: move the clock to T = t_last + 1 before comparing
        shifted = ref.random_projections[j].data.numpy().astype(np.float64) * np.power(np.exp(-lam * 1.0), j)
        np.testing.assert_allclose(shifted, brute[j], rtol=1e-5, atol=1e-5)
        shifted = orc.P[j].astype(np.float64) * np.power(np.exp(-lam * 1.0), j)
        np.testing.assert_allclose(shifted, brute[j], rtol=1e-5, atol=1e-5)
    np.savez_compressed(os.path.join(HERE, 'matrix_kat_brute.npz'), src=s, dst=d, t=t,
                        **{f'A{j}': brute[j] for j in range(L + 1)})
    print('matrix_kat: reference and oracle match the exhaustive walk enumeration')


def sampler_fixture():
    """`recent` NeighborSampler (utils/utils.py:70-224) vs oracle/neighbor_sampler.py."""
    from utils.DataLoader import Data                     # the reference's containers
    from utils.utils import get_neighbor_sampler

    from oracle.neighbor_sampler import RecentNeighborOracle
    rng = np.random.default_rng(77)
    N, E = 37, 400
    src = 1 + (rng.zipf(1.4, E) - 1) % (N - 1)
    dst = 1 + (rng.zipf(1.4, E) - 1) % (N - 1)
    src[5], dst[5] = 9, 9                                  # a self loop
    src[7], dst[7] = src[6], dst[6]                        # a repeated edge
    t = np.sort(np.floor(rng.random(E) * 60.0))            # many equal timestamps
    eid = np.arange(1, E + 1)
    data = Data(src_node_ids=src.astype(np.longlong), dst_node_ids=dst.astype(np.longlong),
                node_interact_times=t.astype(np.float64), edge_ids=eid.astype(np.longlong), labels=np.zeros(E))
    ref = get_neighbor_sampler(data=data, sample_neighbor_strategy='recent', seed=0)
    orc = RecentNeighborOracle(src, dst, eid, t, N)
    out = {'src': src.astype(np.int64), 'dst': dst.astype(np.int64), 'eid': eid.astype(np.int64), 't': t, 'num_nodes': np.int64(N)}
    for k, K in enumerate((1, 5, 20)):
        qn = rng.integers(0, N, 300).astype(np.int64)      # includes node 0 (padding: no history)
        qt = np.where(rng.random(300) < 0.5, np.floor(rng.random(300) * 62.0), rng.random(300) * 62.0)
        a = ref.get_historical_neighbors(qn, qt, num_neighbors=K)
        b = orc.get_historical_neighbors(qn, qt, num_neighbors=K)
        for x, y in zip(a, b):
            assert x.shape == y.shape and np.array_equal(x, y), f'sampler oracle != reference (K={K})'
        out[f'q{k}_nodes'], out[f'q{k}_times'], out[f'q{k}_K'] = qn, qt, np.int64(K)
        out[f'q{k}_nbr'], out[f'q{k}_eid'], out[f'q{k}_t'] = (np.asarray(v) for v in a)
    np.savez_compressed(os.path.join(HERE, 'sampler_tiny.npz'), **out)
    print('sampler_tiny: oracle equals the reference NeighborSampler (recent) for K = 1, 5, 20')


if __name__ == '__main__':
    main()
    sampler_fixture()
