"""GPU suite (-m gpu): the CUDA path, called through the drop-in module and the raw
C ABI, against the oracle (oracle/) and the committed reference fixtures.

Tolerances (north_star: fp32 rtol 1e-5):
  * projections after `update`: rtol 1e-5, atol 1e-6.  Eager decay is bit-exact whenever the
    edge weights are exactly representable (equal timestamps in a batch -> w == 1), because
    everything else follows the reference's rounding points and summation order.  Lazy decay
    applies the product of the skipped per-update factors with ONE rounding (the reference
    rounds once per update): bit-exact when the clock does not move or no update is skipped,
    otherwise within (skipped + 1) * 2^-24 relative — asserted at rtol 1e-5.
  * pair-wise features: |err| <= 1e-5*|ref| + 2e-6*||x_r||*||x_c|| / (1 + max(G,0)) —
    rtol 1e-5 plus the Cauchy-Schwarz floor any fp32 dot product of those rows has
    (the reference's own batched GEMM is only reproducible to that floor).
"""
import ctypes

import numpy as np
import pytest
import torch

from golden_util import CASES, load_case, oracle_kwargs
from oracle.walk_projection import WalkProjectionOracle
from tpnet_b200 import RandomProjectionModule, _lib
import tpnet_b200.random_projection as rpmod

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
MODES = ['eager', 'lazy']


def module_from_cfg(kw, p0, mode):
    m = RandomProjectionModule(device=DEV, beginning_time=np.float64(kw['beginning_time']), decay_mode=mode,
                               **{k: v for k, v in kw.items() if k != 'beginning_time'})
    if not kw['use_matrix']:
        m.random_projections[0].data.copy_(torch.from_numpy(p0))
    return m.to(DEV)


def layers(m):
    m.materialize()
    return [p.data.cpu().numpy() for p in m.random_projections]


def assert_layers(got, ref, mode, exact=True, first=0):
    """eager (and lazy with nothing skipped): bit for bit; lazy: rtol 1e-5 (one rounding per row
    read instead of one per update)."""
    for i in range(first, len(ref)):
        if exact and mode == 'eager':
            assert np.array_equal(got[i], ref[i]), f'layer {i} not bit-exact'
        else:
            scale = max(float(np.abs(ref[i]).max()), 1.0)
            np.testing.assert_allclose(got[i], ref[i], rtol=1e-5, atol=1e-6 * scale, err_msg=f'layer {i}')


def pair_tol(oracle, a, b, ref_raw):
    scale = oracle.pair_norm_bound(a, b)
    return scale / (1.0 + np.maximum(ref_raw, 0))


# --------------------------------------------------------------------------- fixtures of the reference
@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', CASES)
def test_update_matches_reference_fixture(name, mode):
    z, cfg, batches = load_case(name)
    kw = oracle_kwargs(cfg)
    m = module_from_cfg(kw, z['p0'], mode)
    L = kw['num_layer']
    exact = all(np.all(w == 1.0) for _, _, _, w in batches)
    for b, (s, d, t, w) in enumerate(batches):
        m.update(s, d, t)
        if b == 0:
            got = layers(m)
            for i in range(1, L + 1):
                np.testing.assert_allclose(got[i], z[f'after0_P{i}'], rtol=1e-5, atol=1e-6)
            for i in range(2, L + 1):
                assert not got[i].any(), 'layer i must consume the PRE-batch layer i-1 (TPNet.py:90)'
    got = layers(m)
    assert np.array_equal(got[0], z['p0']), 'P_0 is never written by update'
    assert_layers(got, [z['p0']] + [z[f'final_P{i}'] for i in range(1, L + 1)], mode, exact=exact, first=1)
    assert float(m.now_time.item()) == float(z['final_now'])
    m.check_errors()


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', CASES)
def test_pairwise_and_gather_match_reference_fixture(name, mode):
    z, cfg, batches = load_case(name)
    kw = oracle_kwargs(cfg)
    m = module_from_cfg(kw, z['p0'], mode)
    o = WalkProjectionOracle(p0=None if kw['use_matrix'] else z['p0'], **kw)
    for s, d, t, w in batches:
        m.update(s, d, t)
        o.update(s, d, t, weights=w)
    a, b, ref = z['pair_a'], z['pair_b'], z['pair_feat']
    got = m.pair_wise_gram(a, b).cpu().numpy()
    assert got.shape == ref.shape
    o.not_scale = True
    raw = o.pair_wise_gram(a, b, exact=True)
    tol = 1e-5 * np.abs(ref) + 2e-6 * (o.pair_norm_bound(a, b) if kw['not_scale'] else pair_tol(o, a, b, raw)) + 1e-7
    assert np.all(np.abs(got - ref) <= tol), float(np.max(np.abs(got - ref) - tol))
    rows = m.get_random_projections(a)
    assert len(rows) == kw['num_layer'] + 1
    mine = layers(m)
    for i, r in enumerate(rows):
        assert r.shape == (len(a), m.dim)
        assert np.array_equal(r.cpu().numpy(), mine[i][a]), 'gather must return the current rows bit-exactly'
    with torch.no_grad():
        full = m.get_pair_wise_feature(a, b)
    assert full.shape == (len(a), m.pair_wise_feature_dim)
    with torch.no_grad():
        head = m.mlp(torch.from_numpy(got).to(DEV))
    # the no-grad call runs the fused fp32 head (other summation order than cuBLAS): compare at the output's scale
    assert float((full - head).abs().max()) <= 2e-5 * (float(head.abs().max()) + 1.0)


def nbr_tol(o, nbr, src, dst, ref, not_scale):
    """Per-element tolerance of the [m, K, 2F] neighbour features (same rule as pair_tol)."""
    m, k = nbr.shape
    a, b = o.neighbor_pair_lists(nbr, src, dst)
    keep = o.not_scale
    o.not_scale = True
    raw = o.pair_wise_gram(a, b, exact=True)
    o.not_scale = keep
    scale = o.pair_norm_bound(a, b) if not_scale else pair_tol(o, a, b, raw)
    scale = np.concatenate([scale[:m * k], scale[m * k:]], axis=1).reshape(m, k, -1)
    return 1e-5 * np.abs(ref) + 2e-6 * scale + 1e-7


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', CASES)
def test_neighbor_pairwise_matches_reference_fixture(name, mode):
    """tpn_pairwise_neighbors vs the reference's TPNet.py:313-324 output (fixture)."""
    z, cfg, batches = load_case(name)
    kw = oracle_kwargs(cfg)
    m = module_from_cfg(kw, z['p0'], mode)
    o = WalkProjectionOracle(p0=None if kw['use_matrix'] else z['p0'], **kw)
    for s, d, t, w in batches:
        m.update(s, d, t)
        o.update(s, d, t, weights=w)
    nbr, src, dst, ref = z['nbr'], z['nbr_src'], z['nbr_dst'], z['nbr_feat']
    rows, k = nbr.shape
    F = m.pair_wise_feature_dim
    got = m.neighbor_pair_wise_gram(nbr, src, dst)
    assert got.shape == (rows, k, 2, F)
    got = got.cpu().numpy().reshape(rows, k, 2 * F)
    tol = nbr_tol(o, nbr, src, dst, ref, kw['not_scale'])
    assert np.all(np.abs(got - ref) <= tol), float(np.max(np.abs(got - ref) - tol))
    with torch.no_grad():
        full = m.get_neighbor_pair_wise_feature(nbr, src, dst)
        a, b = o.neighbor_pair_lists(nbr, src, dst)
        generic = m.get_pair_wise_feature(a, b)                      # the reference's own call sequence
        generic = torch.cat([generic[:rows * k], generic[rows * k:]], dim=1).reshape(rows, k, -1)
    assert full.shape == (rows, k, 2 * F)
    # same head on the same blocks (fp32 GEMM over a different batch shape: compare at the output's scale)
    assert float((full - generic).abs().max()) <= 1e-5 * max(float(generic.abs().max()), 1.0)
    m.check_errors()


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('N,dim,L,rows,K', [(3000, 140, 3, 400, 20), (500, 120, 2, 37, 7), (800, 210, 3, 64, 33),
                                            (300, 64, 1, 50, 1), (200, 36, 4, 21, 10)])
def test_neighbor_pairwise_vs_oracle(N, dim, L, rows, K, mode):
    """TPNet batch shape (2B rows x K neighbours) and ragged K (not a multiple of 4, more than one
    pass per warp) against the oracle; device-resident ids; structural identities."""
    rng = np.random.default_rng(5)
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=1e-5, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    o = WalkProjectionOracle(**kw)
    m = module_from_cfg(kw, o.P[0], mode)
    for s, d, t in stream(rng, N, 600, 5, 1.25):
        o.update(s, d, t)
        m.update(s, d, t)
    nbr = rng.integers(0, N, (rows, K)).astype(np.int64)
    nbr[0, :] = 0                                             # an all-padding row
    nbr[1, :K // 2] = 0                                       # front padding, as the `recent` sampler pads
    src = rng.integers(1, N, rows).astype(np.int64)
    dst = rng.integers(1, N, rows).astype(np.int64)
    dst[2] = src[2]
    nbr[3, -1] = src[3]
    F = (2 * L + 2) ** 2
    ref = o.neighbor_pair_wise_gram(nbr, src, dst)
    got = m.neighbor_pair_wise_gram(nbr, src, dst).cpu().numpy().reshape(rows, K, 2 * F)
    tol = nbr_tol(o, nbr, src, dst, ref, False)
    assert np.all(np.abs(got - ref) <= tol), float(np.max(np.abs(got - ref) - tol))
    # device-resident ids give the same bits
    dev_ids = [torch.from_numpy(x).to(DEV) for x in (nbr, src, dst)]
    again = m.neighbor_pair_wise_gram(*dev_ids).cpu().numpy().reshape(rows, K, 2 * F)
    assert np.array_equal(again, got)
    # structure: every block is symmetric; the W.W corner is shared by both blocks of (n, k); the
    # S.S / D.D corner is shared by all K neighbours of a row; src == dst gives two equal blocks
    g = got.reshape(rows, K, 2, 2 * L + 2, 2 * L + 2)
    H = L + 1
    assert np.array_equal(g, g.transpose(0, 1, 2, 4, 3))
    assert np.array_equal(g[:, :, 0, :H, :H], g[:, :, 1, :H, :H])
    assert np.array_equal(g[:, :, :, H:, H:], np.broadcast_to(g[:, :1, :, H:, H:], g[:, :, :, H:, H:].shape))
    assert np.array_equal(g[2, :, 0], g[2, :, 1])
    # raw (not_scale) features
    m.not_scale = True
    o.not_scale = True
    raw_ref = o.neighbor_pair_wise_gram(nbr, src, dst)
    raw = m.neighbor_pair_wise_gram(nbr, src, dst).cpu().numpy().reshape(rows, K, 2 * F)
    tolr = nbr_tol(o, nbr, src, dst, raw_ref, True)
    assert np.all(np.abs(raw - raw_ref) <= tolr)
    # empty inputs
    assert m.neighbor_pair_wise_gram(np.zeros((0, K), np.int64), src[:0], dst[:0]).shape == (0, K, 2, F)
    with pytest.raises(IndexError):
        bad = nbr.copy()
        bad[5, 0] = N
        m.neighbor_pair_wise_gram(bad, src, dst)
    m.check_errors()


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', ['wiki_tiny', 'flights_tiny'])
def test_backup_reload_match_reference_fixture(name, mode):
    z, cfg, batches = load_case(name)
    kw = oracle_kwargs(cfg)
    L = kw['num_layer']
    m = module_from_cfg(kw, z['p0'], mode)
    saved = None
    for s, d, t, w in batches:
        m.update(s, d, t)
        if saved is None and float(m.now_time.item()) == float(z['backup_now']):
            cand = m.backup_random_projections()
            if all(np.allclose(cand[1][i].cpu().numpy(), z[f'backup_P{i + 1}'], rtol=1e-5, atol=1e-6)
                   for i in range(L)):
                saved = cand
    assert saved is not None and len(saved[1]) == L and saved[1][0].shape == (kw['node_num'], m.dim)
    assert saved[0].dtype == torch.float64
    m.reload_random_projections(saved)
    assert float(m.now_time.item()) == float(z['backup_now'])
    s, d, t, _ = batches[-1]
    m.update(s, d, t)
    got = layers(m)
    for i in range(1, L + 1):
        np.testing.assert_allclose(got[i], z[f'after_reload_P{i}'], rtol=1e-5, atol=1e-6)


# --------------------------------------------------------------------------- oracle, seeded streams
def stream(rng, N, B, nb, skew, t0=0.0, span=500.0, equal_times=False, frozen=False):
    t = t0
    for _ in range(nb):
        s = 1 + (rng.zipf(skew, B) - 1) % (N - 1)
        d = 1 + (rng.zipf(skew, B) - 1) % (N - 1)
        if equal_times:
            t = t + (0.0 if frozen else float(rng.integers(0, 3)) * 86400.0)
            ts = np.full(B, t)
        else:
            ts = np.sort(t + rng.random(B) * span)
            t = ts[-1]
        yield s.astype(np.int64), d.astype(np.int64), ts.astype(np.float64)


@pytest.mark.parametrize('mode', ['eager', 'lazy', 'lazy-frozen'])
@pytest.mark.parametrize('B,N,dim,L', [(200, 300, 20, 2), (3000, 500, 24, 3), (2500, 4000, 150, 3), (777, 90, 7, 4),
                                       (1, 10, 4, 1), (512, 64, 12, 3), (40000, 5000, 24, 3), (33000, 700, 40, 1)])
def test_update_bit_exact_with_equal_timestamps(B, N, dim, L, mode):
    """With one timestamp per batch every w_j is exactly 1, so the CUDA path must equal the
    oracle bit for bit: decay chain, stable sort-by-target (bitonic for 2B<=4096, radix
    above), batch-order sequential sums, top-down layers.  Sizes cover every code path:
    rank sort (2B<=1024), bitonic (<=4096), radix + short-segment / hub walkers above that.
    'lazy-frozen': the clock never moves, so no decay epoch is created and lazy must be bit-exact
    too (covers the LAZY kernel instantiations bit for bit); 'lazy' with a moving clock is
    checked at rtol 1e-5."""
    frozen = mode == 'lazy-frozen'
    mode = 'lazy' if frozen else mode
    rng = np.random.default_rng(B + N)
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=2e-6, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    o = WalkProjectionOracle(**kw)
    m = module_from_cfg(kw, o.P[0], mode)
    for s, d, t in stream(rng, N, B, 7, 1.3, equal_times=True, frozen=frozen):
        o.update(s, d, t)
        m.update(s, d, t)
    assert_layers(layers(m), o.P, 'eager' if frozen else mode)
    m.check_errors()


@pytest.mark.parametrize('flags', [0, 1, 2, 32, 64], ids=['concurrent', 'per-layer', 'serial', 'no-stream', 'stream-all'])
@pytest.mark.parametrize('mode', ['eager', 'lazy', 'lazy-frozen'])
@pytest.mark.parametrize('B,N,dim,L,skew', [(6000, 300, 210, 3, 1.3), (2100, 50, 140, 3, 1.1), (9000, 2000, 36, 4, 1.5),
                                            (5000, 40, 300, 1, 1.2), (3000, 30, 64, 2, 1.05), (4000, 60, 1000, 2, 1.3),
                                            (3000, 70000, 16, 2, 1.2),       # 1, 2 and 3 radix passes
                                            (20000, 100, 48, 3, 1.4),        # several giants
                                            (4000, 60, 64, 2, 2.2)])         # one target holds 2/3 of the batch: streamed by default
def test_hub_walker_bit_exact(B, N, dim, L, skew, mode, flags):
    """Long segments (>= 64 messages on one target) leave the warp walker for the CTA-pipelined
    hub walkers (cp.async / TMA rings + mbarriers): giant (>= 2048) and regular hubs — giants STREAMED by default
    when one of them holds more than 1/4 of the batch's messages (products materialised by producer CTAs, chains fed by
    bulk copies; 'no-stream': never, 'stream-all': every giant) — 64-float column slices (d=210 -> 216-float rows -> 4 slices of 56/56/56/48 per row; d=1000 -> 16),
    the snapshot + all-layer launches (hub walker on the side stream concurrently with the
    short-segment walker, or serially) and the per-layer launches, eager and lazy decay.
    Equal timestamps make w == 1, so eager (and lazy with a frozen clock) must equal the
    oracle bit for bit."""
    frozen = mode == 'lazy-frozen'
    mode = 'lazy' if frozen else mode
    lib = _lib.load()
    rng = np.random.default_rng(B + N + dim)
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=2e-6, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    o = WalkProjectionOracle(**kw)
    m = module_from_cfg(kw, o.P[0], mode)
    old = lib.tpn_set_debug_flags(flags)
    try:
        for s, d, t in stream(rng, N, B, 5, skew, equal_times=True, frozen=frozen):
            o.update(s, d, t)
            m.update(s, d, t)
        got = layers(m)
    finally:
        lib.tpn_set_debug_flags(old)
    assert_layers(got, o.P, 'eager' if frozen else mode)
    m.check_errors()


@pytest.mark.parametrize('chunk', [256, 1024, 2048])
@pytest.mark.parametrize('mode', ['eager', 'lazy', 'lazy-frozen'])
@pytest.mark.parametrize('B,N,dim,L,skew', [(9000, 300, 210, 3, 1.3), (12000, 50, 140, 3, 1.1), (20000, 2000, 36, 4, 1.5),
                                            (7000, 40, 300, 1, 1.2)])
def test_chunked_accumulation_bit_exact_vs_chunked_oracle(B, N, dim, L, skew, mode, chunk):
    """accumulation='chunked' (tpn_state_t::giant_chunk): rows receiving >= 2048 messages in one update are summed
    chunk by chunk, then the chunk sums in order.  The oracle restates exactly that order, so with w == 1 the
    kernel must equal it BIT FOR BIT (eager; lazy with a frozen clock).  Against the REFERENCE's sequential order it
    can only agree up to that order's own rounding error — a sequential fp32 sum of n >= 2048 terms is off by up to
    ~n * 2^-24 of the row's magnitude (measured 2.4e-4 on the bench workload, where the chunked order is within 3e-6
    of the exact result: scripts/accumulation_order_error.py) — hence the loose second bound, and hence
    accumulation='reference' is the default everywhere a north_star parity claim is made."""
    frozen = mode == 'lazy-frozen'
    mode = 'lazy' if frozen else mode
    rng = np.random.default_rng(B + N + dim + chunk)
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=2e-6, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    o = WalkProjectionOracle(**kw)               # chunked order
    r = WalkProjectionOracle(**kw, p0=o.P[0])    # the reference's sequential order
    m = RandomProjectionModule(device=DEV, decay_mode=mode, accumulation='chunked', giant_chunk=chunk,
                               **{**kw, 'beginning_time': np.float64(0.0)})
    m.random_projections[0].data.copy_(torch.from_numpy(o.P[0]))
    m = m.to(DEV)
    giants = 0
    for s, d, t in stream(rng, N, B, 4, skew, equal_times=True, frozen=frozen):
        giants += int((np.bincount(np.concatenate([s, d]), minlength=N) >= 2048).sum())
        o.update(s, d, t, giant_chunk=chunk)
        r.update(s, d, t)
        m.update(s, d, t)
    assert giants > 0, 'the case must contain giant rows'
    got = layers(m)
    assert_layers(got, o.P, 'eager' if frozen else mode)
    for i in range(1, L + 1):
        row_scale = np.abs(r.P[i]).max(axis=1, keepdims=True)
        assert np.all(np.abs(got[i] - r.P[i]) <= 2e-3 * row_scale + 1e-30), f'layer {i} vs the reference order'
    m.check_errors()


def test_chunked_accumulation_weighted_and_no_giants():
    """Real timestamps (w != 1): still bit-identical to the chunked oracle in eager mode (same rounding points);
    and a batch without giant rows is bit-identical to the reference order whatever the setting."""
    rng = np.random.default_rng(5)
    N, dim, L = 400, 210, 3
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=1e-5, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    o = WalkProjectionOracle(**kw)
    m = RandomProjectionModule(device=DEV, decay_mode='eager', accumulation='chunked', giant_chunk=512,
                               **{**kw, 'beginning_time': np.float64(0.0)})
    m.random_projections[0].data.copy_(torch.from_numpy(o.P[0]))
    m = m.to(DEV)
    for s, d, t in stream(rng, N, 9000, 3, 1.25):
        o.update(s, d, t, giant_chunk=512)
        m.update(s, d, t)
    got = layers(m)
    for i in range(1, L + 1):
        assert np.array_equal(got[i], o.P[i]), f'layer {i}'
    rng = np.random.default_rng(6)
    for _ in range(3):                                      # uniform targets: no row reaches 2048 messages
        s = rng.integers(1, N, 5000).astype(np.int64)
        d = rng.integers(1, N, 5000).astype(np.int64)
        t = np.full(5000, o.now_time + 50.0)
        o.update(s, d, t)                                   # reference order
        m.update(s, d, t)
    got = layers(m)
    for i in range(1, L + 1):
        assert np.array_equal(got[i], o.P[i]), f'layer {i} (no giants)'


@pytest.mark.parametrize('mode', MODES)
def test_hub_walker_weighted_vs_oracle(mode):
    """Same, with real timestamps (weights != 1): rtol 1e-5 against the oracle."""
    rng = np.random.default_rng(77)
    N, B, dim, L = 500, 8000, 210, 3
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=1e-5, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    o = WalkProjectionOracle(**kw)
    m = module_from_cfg(kw, o.P[0], mode)
    for s, d, t in stream(rng, N, B, 4, 1.25):
        o.update(s, d, t)
        m.update(s, d, t)
    got = layers(m)
    for i in range(1, L + 1):
        scale = np.abs(o.P[i]).max()
        np.testing.assert_allclose(got[i], o.P[i], rtol=1e-5, atol=1e-6 * max(scale, 1.0))
    if mode == 'eager':
        # kernel and oracle round exp() once from f64, so even with weights != 1 the results are
        # bit-identical: this is what catches an FMA contraction of fadd(acc, fmul(x, w))
        for i in range(1, L + 1):
            assert np.array_equal(got[i], o.P[i]), f'layer {i}: weighted sums not bit-exact'


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('B,N,dim,L,lam', [(200, 1000, 120, 2, 1e-6), (5000, 3000, 140, 3, 1e-5),
                                           (200, 64, 210, 3, 1e-4)])
def test_update_and_pairwise_vs_oracle(B, N, dim, L, lam, mode):
    rng = np.random.default_rng(11)
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=lam, use_matrix=False,
              beginning_time=100.0, not_scale=False, enforce_dim=dim)
    o = WalkProjectionOracle(**kw)
    m = module_from_cfg(kw, o.P[0], mode)
    for s, d, t in stream(rng, N, B, 6, 1.25, t0=100.0):
        o.update(s, d, t)
        m.update(s, d, t)
    got = layers(m)
    for i in range(1, L + 1):
        scale = np.abs(o.P[i]).max()
        np.testing.assert_allclose(got[i], o.P[i], rtol=1e-5, atol=1e-6 * max(scale, 1.0))
        if mode == 'eager':
            assert np.array_equal(got[i], o.P[i]), f'layer {i}: weighted sums not bit-exact'
    n = 4099                                              # ragged: not a multiple of the 16-pair CTA tile
    a = rng.integers(0, N, n).astype(np.int64)
    b = rng.integers(0, N, n).astype(np.int64)
    b[:50] = a[:50]                                       # self pairs
    feat = m.pair_wise_gram(a, b).cpu().numpy()
    ref = o.pair_wise_gram(a, b)
    o.not_scale = True
    raw = o.pair_wise_gram(a, b, exact=True)
    tol = 1e-5 * np.abs(ref) + 2e-6 * pair_tol(o, a, b, raw) + 1e-7
    assert np.all(np.abs(feat - ref) <= tol), float(np.max(np.abs(feat - ref) - tol))
    # raw (not_scale) features and row-major (r, c) ordering / symmetry
    m.not_scale = True
    g = m.pair_wise_gram(a, b).cpu().numpy().reshape(n, 2 * L + 2, 2 * L + 2)
    assert np.array_equal(g, g.transpose(0, 2, 1))
    tolr = 1e-5 * np.abs(raw) + 2e-6 * o.pair_norm_bound(a, b) + 1e-7
    assert np.all(np.abs(g.reshape(n, -1) - raw) <= tolr)
    # swapping the endpoints permutes the blocks exactly
    gs = m.pair_wise_gram(b, a).cpu().numpy().reshape(n, 2 * L + 2, 2 * L + 2)
    H = L + 1
    assert np.array_equal(gs[:, :H, :H], g[:, H:, H:]) and np.array_equal(gs[:, :H, H:], g[:, H:, :H])


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('N,dim,L,n', [(5000, 210, 3, 40003), (3000, 140, 3, 33000), (2000, 120, 2, 36001),
                                       (1500, 64, 1, 32770), (1200, 72, 4, 34000)])
def test_pairwise_large_call(N, dim, L, n, mode):
    """Decoder-sized calls (tens of thousands of pairs, ragged tail): same bits as the same pairs
    encoded in 5,000-pair calls, and the oracle's values."""
    rng = np.random.default_rng(23)
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=2e-5, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    o = WalkProjectionOracle(**kw)
    m = module_from_cfg(kw, o.P[0], mode)
    for s, d, t in stream(rng, N, 1500, 5, 1.25):
        o.update(s, d, t)
        m.update(s, d, t)
    a = rng.integers(0, N, n).astype(np.int64)
    b = rng.integers(0, N, n).astype(np.int64)
    b[:100] = a[:100]
    big = m.pair_wise_gram(a, b).cpu().numpy()
    # >= 5,000-pair calls (same kernel, same 4-pair groups; the remainder rides with the last call)
    cuts = list(range(0, n - 5000, 5000)) + [n]
    small = np.concatenate([m.pair_wise_gram(a[i:j], b[i:j]).cpu().numpy() for i, j in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(big, small)
    ref = o.pair_wise_gram(a, b)
    o.not_scale = True
    raw = o.pair_wise_gram(a, b, exact=True)
    tol = 1e-5 * np.abs(ref) + 2e-6 * pair_tol(o, a, b, raw) + 1e-7
    assert np.all(np.abs(big - ref) <= tol), float(np.max(np.abs(big - ref) - tol))
    assert np.all(np.abs(small - ref) <= tol)
    g = big.reshape(n, 2 * L + 2, 2 * L + 2)
    assert np.array_equal(g, g.transpose(0, 2, 1))
    dev_ids = [torch.from_numpy(x).to(DEV) for x in (a, b)]
    assert np.array_equal(m.pair_wise_gram(*dev_ids).cpu().numpy(), big)
    m.check_errors()


@pytest.fixture(params=['ffma', 'tensor'])
def head_kernel(request):
    """Both implementations of tpn_head_forward: packed-FFMA (fp32 CUDA cores) and tcgen05 (fp16 x 2 split operands,
    fp32 accumulation in TMEM; the default), selected by TPN_DEBUG_HEAD_FFMA."""
    lib = _lib.load()
    old = lib.tpn_set_debug_flags(8 if request.param == 'ffma' else 0)
    yield request.param
    lib.tpn_set_debug_flags(old)


@pytest.mark.parametrize('n', [1, 63, 64, 65, 127, 128, 129, 1000, 100003])
def test_fused_head_matches_torch_head(n, head_kernel):
    """tpn_head_forward (no-grad path of get_pair_wise_feature) vs `self.mlp` in PyTorch: both are fp32-accurate
    with different summation orders (the tensor-core kernel: operands split into two fp16 numbers, 22 bits, fp32
    accumulation), so both are compared with the float64 head."""
    torch.manual_seed(3)
    kw = dict(node_num=50, edge_num=5000, dim_factor=10, num_layer=3, time_decay_weight=1e-6, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=-1)
    m = RandomProjectionModule(device=DEV, decay_mode='eager', **kw).to(DEV)
    assert m.pair_wise_feature_dim == 64
    with torch.no_grad():
        for p in m.mlp.parameters():
            p.mul_(3.0)                                       # well away from the init scale
    x = (torch.rand(n, 64, device=DEV) * 12.0).contiguous()   # log-scaled Gram entries live in [0, ~12]
    x[:, ::7] = 0
    with torch.no_grad():
        fused = m._head(x)
        ref32 = m.mlp(x)
    ref64 = m.mlp.double()(x.double())
    m.mlp.float()
    scale = float(ref64.abs().max()) + 1.0
    err_fused = float((fused.double() - ref64).abs().max())
    err_torch = float((ref32.double() - ref64).abs().max())
    assert err_fused <= 4 * err_torch + 1e-6 * scale, (err_fused, err_torch)
    assert float((fused - ref32).abs().max()) <= 2e-5 * scale
    # autograd on: the PyTorch head, gradients reach its parameters (TPNet trains it)
    y = m._head(x)
    assert y.requires_grad and torch.equal(y.detach(), ref32)
    # the public calls use it: same values with the fused head switched off
    ids = np.arange(1, 41, dtype=np.int64)
    with torch.no_grad():
        a = m.get_pair_wise_feature(ids, ids[::-1].copy())
        m.fused_head = False
        b = m.get_pair_wise_feature(ids, ids[::-1].copy())
    assert float((a - b).abs().max()) <= 2e-5 * (float(b.abs().max()) + 1.0)


def test_tensor_core_head_dynamic_range_and_counts(head_kernel):
    """Rows of very different magnitude (each row of the activations has its own power-of-two scale), raw Gram values
    (`not_scale`: 1e6 and beyond — far outside fp16's range before scaling), all-zero rows, tiny weights; and the
    device-side row count of routed calls (rows past the count stay untouched)."""
    torch.manual_seed(5)
    mlp = torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.ReLU(), torch.nn.Linear(256, 64)).to(DEV)
    lib = _lib.load()
    n = 777
    x = torch.rand(n, 64, device=DEV)
    x *= torch.logspace(-6, 7, n, device=DEV)[:, None]            # row magnitudes from 1e-6 to 1e7
    x[5] = 0
    x[::9, ::3] *= -1.0
    for wscale in (1.0, 1e-4, 300.0):
        with torch.no_grad():
            for p in mlp.parameters():
                p.mul_(wscale)
        ref64 = mlp.double()(x.double())
        mlp.float()
        y = torch.full((n, 64), 7.0, device=DEV)
        cnt = torch.tensor([n - 100], dtype=torch.int32, device=DEV)
        l1, l2 = mlp[0], mlp[2]
        rc = lib.tpn_head_forward(x.data_ptr(), n, cnt.data_ptr(), 64, 256, l1.weight.data_ptr(), l1.bias.data_ptr(),
                                  l2.weight.data_ptr(), l2.bias.data_ptr(), y.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        assert torch.all(y[n - 100:] == 7.0)                                   # rows past the device-side count
        got, want = y[:n - 100].double(), ref64[:n - 100]
        with torch.no_grad():
            ref32 = mlp(x)[:n - 100].double()
        # per row: error relative to the row's largest output (fp32 SGEMM noise is relative to sum |x||w|, so the
        # torch fp32 result itself is only this accurate)
        row = want.abs().max(dim=1, keepdim=True).values + 1e-30
        err = ((got - want).abs() / row).max().item()
        err32 = ((ref32 - want).abs() / row).max().item()
        assert err <= 4 * err32 + 2e-6, (head_kernel, wscale, err, err32)
        with torch.no_grad():
            for p in mlp.parameters():
                p.div_(wscale)


def test_lazy_matches_eager_and_log_restart(monkeypatch):
    """Lazy decay = one multiply by the product of the skipped factors.  A tiny log forces
    materialise + restart several times; readers rescale on the fly (no materialise)."""
    monkeypatch.setattr(rpmod, '_DEFAULT_LOG_EPOCHS', 5)
    rng = np.random.default_rng(5)
    N, L, dim = 800, 3, 36
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=3e-4, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    torch.manual_seed(0)
    e = RandomProjectionModule(device=DEV, decay_mode='eager', **kw).to(DEV)
    z = RandomProjectionModule(device=DEV, decay_mode='lazy', **kw).to(DEV)
    z.random_projections[0].data.copy_(e.random_projections[0].data)
    ids = rng.integers(0, N, 512).astype(np.int64)
    ids2 = rng.integers(0, N, 512).astype(np.int64)
    for k, (s, d, t) in enumerate(stream(rng, N, 150, 23, 1.2)):
        e.update(s, d, t)
        z.update(s, d, t)
        if k % 5 == 4:     # on-the-fly rescale inside the readers, no materialise
            ge, gz = e.pair_wise_gram(ids, ids2), z.pair_wise_gram(ids, ids2)
            assert torch.allclose(ge, gz, rtol=1e-5, atol=1e-5)
            for x, y in zip(e.get_random_projections(ids), z.get_random_projections(ids)):
                assert torch.allclose(x, y, rtol=1e-5, atol=1e-7)
    assert_layers(layers(z), layers(e), 'lazy')
    sd = z.state_dict()        # materialises
    assert torch.allclose(sd['random_projections.2'], e.random_projections[2].data, rtol=1e-5, atol=1e-7)


def test_lazy_long_gaps_and_underflow():
    """Rows untouched for hundreds of updates: the single-multiply rescale stays within
    (gap + 1) * 2^-24 of the reference's per-update chain (asserted: rtol 1e-5).  Then a decay
    so strong that the fp32 factors underflow: the f64 product log is restarted, never NaN."""
    rng = np.random.default_rng(8)
    N, L, dim = 5000, 3, 24
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=2e-5, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    torch.manual_seed(1)
    e = RandomProjectionModule(device=DEV, decay_mode='eager', **kw).to(DEV)
    z = RandomProjectionModule(device=DEV, decay_mode='lazy', **kw).to(DEV)
    z.random_projections[0].data.copy_(e.random_projections[0].data)
    s0 = np.arange(1, N, dtype=np.int64)
    e.update(s0, s0[::-1].copy(), np.zeros(N - 1))                      # touch every row once
    z.update(s0, s0[::-1].copy(), np.zeros(N - 1))
    for s, d, t in stream(rng, N, 8, 400, 1.3, t0=0.0, span=50.0):      # 400 updates, 16 rows each
        e.update(s, d, t)
        z.update(s, d, t)
    ge, gz = layers(e), layers(z)
    for i in range(1, L + 1):
        row_scale = np.abs(ge[i]).max(axis=1, keepdims=True)             # sums of messages can cancel
        err = np.abs(gz[i] - ge[i])
        assert np.all(err <= 1e-5 * np.abs(ge[i]) + 1e-6 * row_scale), float((err / (row_scale + 1e-30)).max())
        assert ge[i].any()
    # underflow: lambda * dt ~ 50 per update -> c_3 = exp(-150) = 0 in fp32
    t = float(e.now_time.item())
    for k in range(6):
        s = rng.integers(1, N, 64).astype(np.int64)
        d = rng.integers(1, N, 64).astype(np.int64)
        ts = np.full(64, t + 2.5e6 * (k + 1))
        e.update(s, d, ts)
        z.update(s, d, ts)
    ge, gz = layers(e), layers(z)
    for i in range(1, L + 1):
        assert np.isfinite(gz[i]).all()
        np.testing.assert_allclose(gz[i], ge[i], rtol=1e-5, atol=1e-30)


def test_reset_and_rng_parity():
    kw = dict(node_num=50, edge_num=500, dim_factor=2, num_layer=2, time_decay_weight=1e-3, use_matrix=False,
              beginning_time=np.float64(3.0), not_scale=False, enforce_dim=-1)
    for mode in MODES:
        m = RandomProjectionModule(device=DEV, decay_mode=mode, **kw).to(DEV)
        ids = np.arange(1, 40, dtype=np.int64)
        m.update(ids, ids[::-1].copy(), np.linspace(4.0, 9.0, len(ids)))
        assert layers(m)[1].any()
        torch.manual_seed(123)
        m.reset_random_projections()
        got = layers(m)
        assert not got[1].any() and not got[2].any()
        assert float(m.now_time.item()) == 3.0 and m._now_host == 3.0
        # same draw as nn.init.normal_ on a contiguous [N, d] CUDA parameter (TPNet.py:139)
        torch.manual_seed(123)
        expect = torch.empty(50, m.dim, device=DEV).normal_(0, 1 / np.sqrt(m.dim))
        assert torch.equal(m.random_projections[0].data, expect)
        assert not m._state[:, :, m.dim:].any()


def test_edge_cases_and_errors():
    kw = dict(node_num=20, edge_num=100, dim_factor=1, num_layer=2, time_decay_weight=1e-3, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=6)
    m = RandomProjectionModule(device=DEV, **kw).to(DEV)
    e = np.array([], dtype=np.int64)
    assert m.get_pair_wise_feature(e, e).shape == (0, 36)
    assert [r.shape for r in m.get_random_projections(e)] == [(0, 6)] * 3
    with pytest.raises(IndexError):
        m.update(e, e, np.array([], dtype=np.float64))
    with pytest.raises(IndexError):
        m.update(np.array([1]), np.array([20]), np.array([1.0]))
    with pytest.raises(IndexError):
        m.update(np.array([-1]), np.array([2]), np.array([1.0]))
    with pytest.raises(IndexError):
        m.get_pair_wise_feature(np.array([25]), np.array([1]))
    # device-resident ids: out-of-range edges are dropped and flagged
    s = torch.tensor([1, 99], device=DEV); d = torch.tensor([2, 3], device=DEV)
    t = torch.tensor([1.0, 2.0], device=DEV, dtype=torch.float64)
    m.update(s, d, t, next_time=2.0)
    with pytest.raises(IndexError):
        m.check_errors()
    got = layers(m)
    assert got[1][1].any() and got[1][2].any() and not got[1][3].any()
    # int32 ids and python lists behave like int64 arrays
    f1 = m.pair_wise_gram(np.array([1, 2], dtype=np.int32), [2, 1])
    f2 = m.pair_wise_gram(np.array([1, 2]), np.array([2, 1]))
    assert torch.equal(f1, f2)


def test_use_matrix_wide_rows_multi_column_tiles():
    """use_matrix=True with N > 1024 exercises the column-tiled walk kernel (row wider
    than one 32-lane x 8-float4 tile) in both decay modes."""
    rng = np.random.default_rng(2)
    N, L = 1100, 2
    kw = dict(node_num=N, edge_num=4000, dim_factor=1, num_layer=L, time_decay_weight=1e-5, use_matrix=True,
              beginning_time=0.0, not_scale=True, enforce_dim=-1)
    o = WalkProjectionOracle(**kw)
    ms = [module_from_cfg(kw, None, mode) for mode in MODES]
    for s, d, t in stream(rng, N, 64, 4, 1.5, equal_times=True):
        o.update(s, d, t)
        for m in ms:
            m.update(s, d, t)
    for m, mode in zip(ms, MODES):
        assert_layers(layers(m), o.P, mode)


def test_c_abi_direct_with_padded_node_stride():
    """Raw C-ABI call (no module): node_stride larger than (L+1)*row_stride, error codes."""
    lib = _lib.load()
    N, L, d, rs, ns = 33, 2, 10, 12, 48
    rng = np.random.default_rng(0)
    host = np.zeros((N, ns), dtype=np.float32)
    for l in range(L + 1):
        host[:, l * rs:l * rs + d] = rng.standard_normal((N, d)).astype(np.float32)
    state = torch.from_numpy(host).to(DEV)
    st = _lib.TpnState(data=state.data_ptr(), num_nodes=N, num_layer=L, dim=d, row_stride=rs, node_stride=ns,
                       stamps=None, decay_log=None, log_capacity=0, epoch=0, cum_floor=1.0)
    a = torch.from_numpy(rng.integers(0, N, 21)).to(DEV)
    b = torch.from_numpy(rng.integers(0, N, 21)).to(DEV)
    out = torch.empty(21, 36, device=DEV)
    stream_ptr = torch.cuda.current_stream().cuda_stream
    assert lib.tpn_pairwise(ctypes.byref(st), a.data_ptr(), b.data_ptr(), 21, None, 0, out.data_ptr(), stream_ptr) == 0
    x = np.concatenate([host[a.cpu().numpy()].reshape(21, 4, rs)[:, :3, :d],
                        host[b.cpu().numpy()].reshape(21, 4, rs)[:, :3, :d]], axis=1).astype(np.float64)
    ref = np.einsum('nrd,ncd->nrc', x, x).reshape(21, 36)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    g = torch.empty(3, 21, d, device=DEV)
    assert lib.tpn_gather(ctypes.byref(st), a.data_ptr(), 21, g.data_ptr(), stream_ptr) == 0
    assert np.array_equal(g[1].cpu().numpy(), host[a.cpu().numpy(), rs:rs + d])
    st.row_stride = 10                                   # not a multiple of 4 floats
    assert lib.tpn_pairwise(ctypes.byref(st), a.data_ptr(), b.data_ptr(), 21, None, 0, out.data_ptr(), stream_ptr) == -1
    st.row_stride = rs
    ws = torch.empty(64, dtype=torch.uint8, device=DEV)
    t = torch.zeros(21, dtype=torch.float64, device=DEV)
    rc = lib.tpn_update(ctypes.byref(st), a.data_ptr(), b.data_ptr(), t.data_ptr(), 21, 0.0, -1e-3, None,
                        ws.data_ptr(), 64, None, stream_ptr)
    assert rc == -2                                      # workspace too small
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.tpn_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)) == 0
    assert sm.value > 0 and major.value >= 10, 'this library is built for sm_100a only'


# --------------------------------------------------------------------------- BASELINE sizes: properties
@pytest.mark.parametrize('mode', MODES)
def test_full_size_reddit_shape_properties(mode):
    """Reddit-shaped state (N=10,985, d=140, L=3), B=200, and one 100k-edge batch through the
    radix path: size-independent properties instead of the (too slow) oracle."""
    rng = np.random.default_rng(9)
    N, L, dim = 10985, 3, 140
    kw = dict(node_num=N, edge_num=672448, dim_factor=10, num_layer=L, time_decay_weight=1e-6, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=-1)
    torch.manual_seed(0)
    m = RandomProjectionModule(device=DEV, decay_mode=mode, **kw).to(DEV)
    assert m.dim == dim
    before = [x.astype(np.float64) for x in layers(m)]
    # (1) conservation: with w == 1 and no clock movement, sum_u P_1[u] grows by sum_j (P_0[src_j] + P_0[dst_j])
    s = (1 + (rng.zipf(1.2, 100000) - 1) % (N - 1)).astype(np.int64)
    d = (1 + (rng.zipf(1.2, 100000) - 1) % (N - 1)).astype(np.int64)
    t = np.zeros(100000)
    m.update(s, d, t)
    after = [x.astype(np.float64) for x in layers(m)]
    grow = after[1].sum(0) - before[1].sum(0)
    expect = before[0][s].sum(0) + before[0][d].sum(0)
    np.testing.assert_allclose(grow, expect, rtol=1e-4, atol=1e-2)
    assert not after[2].any() and np.array_equal(after[0], before[0])
    # (2) against a plain PyTorch fp32 reference of the same op on the GPU (index_add_, any order)
    P0 = torch.from_numpy(before[0]).float().to(DEV)
    ref1 = torch.zeros_like(P0)
    sd_, dd_ = torch.from_numpy(s).to(DEV), torch.from_numpy(d).to(DEV)
    ref1.index_add_(0, sd_, P0[dd_])
    ref1.index_add_(0, dd_, P0[sd_])
    scale = float(ref1.abs().max())
    np.testing.assert_allclose(after[1], ref1.cpu().numpy(), rtol=1e-4, atol=2e-5 * scale)
    # (3) decoder + encoder shaped pair batches stay finite, non-negative, symmetric
    for s, d, t in stream(rng, N, 200, 5, 1.2, t0=1.0, span=4000.0):
        m.update(s, d, t)
    K = 20
    nbr = rng.integers(0, N, (400, K))
    a = np.tile(nbr.reshape(-1), 2).astype(np.int64)
    b = np.concatenate([np.repeat(np.tile(s, 2), K), np.repeat(np.tile(d, 2), K)]).astype(np.int64)
    f = m.pair_wise_gram(a, b)
    assert f.shape == (16000, 64) and bool(torch.isfinite(f).all()) and bool((f >= 0).all())
    g = f.reshape(-1, 8, 8)
    assert torch.equal(g, g.transpose(1, 2))
    m.check_errors()


# --------------------------------------------------------------------------- BASELINE shapes against the oracle
@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', ['wikipedia', 'reddit', 'flights'])
def test_baseline_shape_stream_vs_oracle(name, mode):
    """BASELINE.json configs[0..2] at FULL shape (Wikipedia 9,228 x d=120 L=2; Reddit 10,985 x d=140 L=3; Flights 13,170 x
    d=150 L=3, day-quantised timestamps): a stream of TPNet batches (B=200) from the bench's own generator, the oracle run
    beside the CUDA path.  Eager decay: BIT-EXACT after every batch (same rounding points, exp rounded once from f64 on
    both sides); lazy: rtol 1e-5.  Then the decoder call (B pairs) and the encoder's structured call (2B rows x K=20
    neighbours = 16,000 pair blocks) against the oracle's features."""
    from tpnet_b200.synth import SHAPES, RecentNeighbors, edge_stream, tpnet_neighbor_batch
    shape = SHAPES[name]
    kw = dict(node_num=shape.node_num, edge_num=shape.edge_num, dim_factor=shape.dim_factor, num_layer=shape.num_layer,
              time_decay_weight=shape.time_decay_weight, use_matrix=False, beginning_time=0.0, not_scale=False,
              enforce_dim=-1)
    o = WalkProjectionOracle(**kw)
    assert o.dim == shape.dim
    m = module_from_cfg(kw, o.P[0], mode)
    nbr = RecentNeighbors(shape.node_num, 20)
    batches = list(edge_stream(shape, 200, 12, seed=7))
    if name == 'flights':
        assert all(len(np.unique(t)) <= 2 for _, _, t in batches)            # day stamps: (almost) all weights are 1
    for i, (s, d, t) in enumerate(batches):
        o.update(s, d, t)
        m.update(s, d, t)
        nbr.insert(s, d)
        if i in (0, 5, 11):
            got = layers(m)
            if mode == 'eager':
                for l in range(1, shape.num_layer + 1):
                    assert np.array_equal(got[l], o.P[l]), (name, i, l)
            else:
                assert_layers(got, o.P, mode)
    s, d, t = batches[-1]
    rng = np.random.default_rng(3)
    neg = rng.integers(1, shape.node_num, 200).astype(np.int64)
    for a, b in ((s, d), (s, neg)):                                          # decoder shape
        got = m.pair_wise_gram(a, b).cpu().numpy()
        ref = o.pair_wise_gram(a, b, exact=True)
        raw = np.expm1(ref.astype(np.float64))
        tol = 1e-5 * np.abs(ref) + 2e-6 * pair_tol(o, a, b, raw)
        assert np.all(np.abs(got - ref) <= tol + 1e-7), name
    nb, s2, d2 = tpnet_neighbor_batch(nbr, s, d)                             # encoder shape: [2B, K] neighbours
    got = m.neighbor_pair_wise_gram(nb, s2, d2).cpu().numpy().reshape(400, 20, -1)
    ref = o.neighbor_pair_wise_gram(nb, s2, d2, exact=True)
    a_ids, b_ids = o.neighbor_pair_lists(nb, s2, d2)
    raw = np.expm1(o.pair_wise_gram(a_ids, b_ids, exact=True).astype(np.float64))
    tol = 1e-5 * np.abs(o.pair_wise_gram(a_ids, b_ids, exact=True)) + 2e-6 * pair_tol(o, a_ids, b_ids, raw)
    F = m.pair_wise_feature_dim
    tol = np.concatenate([tol[:400 * 20], tol[400 * 20:]], axis=1).reshape(400, 20, 2 * F)
    assert np.all(np.abs(got - ref) <= tol + 1e-7), name
    m.check_errors()


@pytest.mark.parametrize('accumulation', ['reference', 'chunked'])
def test_powerlaw_replica_vs_oracle(accumulation):
    """The headline regime, down-scaled in node count only: power-law graph (zipf 1.2), 100,001 nodes,
    d=210, L=3, lambda=1e-7, LAZY decay, batches of 100,000 edges (top hub: ~35,600 messages per batch, radix sort,
    short-segment walker + hub walkers) against the oracle (np.add.at: the reference's sequential order; or the oracle's
    restatement of the chunked order).  rtol 1e-5 element-wise with an absolute floor of 1e-6 of the layer's largest
    entry (lazy decay rounds once per read instead of once per update)."""
    import dataclasses
    from tpnet_b200.synth import SHAPES, edge_stream
    shape = dataclasses.replace(SHAPES['powerlaw'], num_src=100_000)
    kw = dict(node_num=shape.node_num, edge_num=shape.edge_num, dim_factor=shape.dim_factor, num_layer=shape.num_layer,
              time_decay_weight=shape.time_decay_weight, use_matrix=False, beginning_time=0.0, not_scale=False,
              enforce_dim=-1)
    o = WalkProjectionOracle(**kw)
    assert o.dim == 210
    m = RandomProjectionModule(device=DEV, decay_mode='lazy', accumulation=accumulation, giant_chunk=1024,
                               **{**kw, 'beginning_time': np.float64(0.0)})
    m.random_projections[0].data.copy_(torch.from_numpy(o.P[0]))
    m = m.to(DEV)
    chunk = 1024 if accumulation == 'chunked' else 0
    for s, d, t in edge_stream(shape, 100_000, 3, seed=1234):
        assert np.bincount(np.concatenate([s, d])).max() > 30000
        o.update(s, d, t, giant_chunk=chunk)
        m.update(s, d, t)
    got = layers(m)
    for l in range(1, 4):
        scale = float(np.abs(o.P[l]).max())
        np.testing.assert_allclose(got[l], o.P[l], rtol=1e-5, atol=1e-6 * scale, err_msg=f'layer {l}')
    a = np.arange(1, 5001, dtype=np.int64)
    b = np.random.default_rng(0).integers(1, shape.node_num, 5000).astype(np.int64)
    got_f = m.pair_wise_gram(a, b).cpu().numpy()
    ref = o.pair_wise_gram(a, b, exact=True)
    raw = np.expm1(ref.astype(np.float64))
    assert np.all(np.abs(got_f - ref) <= 1e-5 * np.abs(ref) + 2e-6 * pair_tol(o, a, b, raw) + 1e-7)
    m.check_errors()


def test_step_graphs_replay_equals_eager_calls():
    """tpnet_b200.pipeline.StepGraphs: one CUDA graph per batch of a device-resident split (decoder
    features of the positive / pre-drawn negative pairs + update), replayed in order for two "epochs" (reset in between),
    against the same calls issued eagerly: features and final state bit-identical; out-of-order replay is refused."""
    from tpnet_b200.pipeline import EpochBatches, StepGraphs
    rng = np.random.default_rng(12)
    N, E, B = 700, 2050, 200
    src = rng.integers(1, N, E); dst = rng.integers(1, N, E); neg = rng.integers(1, N, E)
    t = np.sort(rng.random(E) * 5e5)
    kw = dict(node_num=N, edge_num=E + 1, dim_factor=10, num_layer=3, time_decay_weight=1e-6, use_matrix=False,
              beginning_time=np.float64(t[0]), not_scale=False, enforce_dim=-1)
    for mode in MODES:
        torch.manual_seed(1)
        m = RandomProjectionModule(device=DEV, decay_mode=mode, **kw).to(DEV)
        torch.manual_seed(1)
        e = RandomProjectionModule(device=DEV, decay_mode=mode, **kw).to(DEV)
        e.load_state_dict(m.state_dict())
        eb = EpochBatches(src, dst, t, B, DEV, extra={'neg': neg})
        sg = StepGraphs(m, list(eb))
        assert len(sg) == 11
        p0 = m.random_projections[0].data.clone()
        for epoch in range(2):
            for b in eb:
                out = sg.replay(b.index)
                with torch.no_grad():
                    want_pos = e.get_pair_wise_feature(b.src, b.dst)
                    want_neg = e.get_pair_wise_feature(b.src, b.extra['neg'])
                    e.update(b.src, b.dst, b.t, next_time=b.t_last)
                assert torch.equal(out['pos'][:len(b)], want_pos) and torch.equal(out['neg'][:len(b)], want_neg), (mode, b.index)
            assert float(m.now_time) == float(e.now_time) == t[-1]
            m.materialize(); e.materialize()
            for i in range(1, 4):
                assert torch.equal(m.random_projections[i].data, e.random_projections[i].data), (mode, epoch, i)
            if epoch == 0:
                with pytest.raises(RuntimeError, match='out of order'):
                    sg.replay(3)
                for mod in (m, e):                               # next epoch: same start state (same P_0: copied back)
                    mod.reset_random_projections()
                    mod.random_projections[0].data.copy_(p0)
                sg.rewind()
        m.check_errors()


@pytest.mark.parametrize('N,dim,L,B', [(40_000, 16, 2, 30_000),          # 16-bit keys: 2 radix passes
                                       (200_001, 24, 3, 70_001),         # 18 bits: 3 passes, ragged last tile
                                       (17_000_000, 8, 1, 50_000),       # 25 bits: 4 passes (both histogram buffers reused)
                                       (3_000, 40, 3, 2_049)])           # just above the single-CTA sort: 3 tiles
def test_fused_front_end_equals_separate_launches(N, dim, L, B):
    """Large batches: the sort front end as ONE cooperative launch with grid barriers (default) against the same
    phases as 12-13 separate launches (TPN_DEBUG_LEGACY_FRONT).  Stable sort, same payload: the states must be
    bit-identical, with hubs (zipf endpoints), equal and real timestamps, lazy and eager decay."""
    lib = _lib.load()
    rng = np.random.default_rng(5)
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=1e-6, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    for mode in (['lazy'] if N > 1_000_000 else MODES):
        torch.manual_seed(3)
        a = RandomProjectionModule(device=DEV, decay_mode=mode, **kw).to(DEV)
        b = RandomProjectionModule(device=DEV, decay_mode=mode, **kw).to(DEV)
        b.random_projections[0].data.copy_(a.random_projections[0].data)
        t = 0.0
        for it in range(3):
            s = (1 + (rng.zipf(1.3, B) - 1) % (N - 1)).astype(np.int64)
            d = rng.integers(1, N, B).astype(np.int64)
            ts = np.sort(t + rng.random(B) * 1000.0) if it else np.full(B, 50.0)
            t = float(ts[-1])
            a.update(s, d, ts)
            old = lib.tpn_set_debug_flags(16)
            try:
                b.update(s, d, ts)
            finally:
                lib.tpn_set_debug_flags(old)
        a.materialize()
        b.materialize()
        for i in range(1, L + 1):
            assert torch.equal(a.random_projections[i].data, b.random_projections[i].data), (mode, i)
        a.check_errors()
        b.check_errors()
        del a, b
        torch.cuda.empty_cache()


@pytest.mark.parametrize('mode', MODES)
def test_update_prepare_overlaps_reads_and_changes_nothing(mode):
    """`update_prepare` starts the half of the update that does not write the state on a side stream; pair-wise calls
    issued between it and `update` still see the PRE-batch state, and the state after `update` is bit-identical to a
    module that never prepared.  Also: a prepared half that is not followed by the matching update (different arrays,
    reset in between) is dropped and the whole update runs."""
    rng = np.random.default_rng(9)
    N, dim, L, B = 5000, 48, 3, 6000
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=1e-6, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    torch.manual_seed(4)
    a = RandomProjectionModule(device=DEV, decay_mode=mode, **kw).to(DEV)
    b = RandomProjectionModule(device=DEV, decay_mode=mode, **kw).to(DEV)
    b.random_projections[0].data.copy_(a.random_projections[0].data)
    t = 0.0
    for it in range(5):
        s = (1 + (rng.zipf(1.4, B) - 1) % (N - 1)).astype(np.int64)
        d = rng.integers(1, N, B).astype(np.int64)
        ts = np.sort(t + rng.random(B) * 5000.0)
        t = float(ts[-1])
        ds, dd, dt = (torch.from_numpy(x).to(DEV) for x in (s, d, ts))
        with torch.no_grad():
            ref_feat = b.pair_wise_gram(ds, dd)
            if it == 3:                                      # prepared, then a DIFFERENT batch arrives: whole update
                assert a.update_prepare(dd, ds, dt, next_time=t)
            else:
                assert a.update_prepare(ds, dd, dt, next_time=t)
            got_feat = a.pair_wise_gram(ds, dd)              # reads the pre-batch state while the side stream works
            assert torch.equal(got_feat, ref_feat), it
        a.update(ds, dd, dt, next_time=t)
        b.update(ds, dd, dt, next_time=t)
        if it == 1:                                          # numpy arrays through the staging ring
            s2 = rng.integers(1, N, B).astype(np.int64)
            ts2 = np.sort(t + rng.random(B) * 10.0)
            t = float(ts2[-1])
            assert a.update_prepare(s2, d, ts2)
            a.update(s2, d, ts2)
            b.update(s2, d, ts2)
    a.materialize()
    b.materialize()
    for i in range(1, L + 1):
        assert torch.equal(a.random_projections[i].data, b.random_projections[i].data), i
    assert a.update_prepare(ds, dd, dt, next_time=t + 1.0)
    a.reset_random_projections()                             # drops the prepared half
    assert a._pending is None
    assert not a.update_prepare(ds[:100], dd[:100], dt[:100], next_time=t)      # single-CTA sort path: nothing to split
    a.check_errors()
    b.check_errors()


def test_feature_stream_calls_equal_sequential_calls():
    """The two decoder calls of a batch on two streams (module.feature_stream(): pipeline.tpnet_step, bench e2e): the
    (src, neg) call runs on the feature stream with its own staging ring while the (src, dst) call and update_prepare
    run on the current / prepare streams.  Same features, same state as a module that does everything in sequence —
    through more calls than a staging ring has slots, with numpy and with device-resident ids."""
    rng = np.random.default_rng(21)
    N, dim, L, B = 4000, 40, 3, 5000
    kw = dict(node_num=N, edge_num=10 * N, dim_factor=1, num_layer=L, time_decay_weight=1e-6, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=dim)
    torch.manual_seed(5)
    a = RandomProjectionModule(device=DEV, decay_mode='lazy', **kw).to(DEV)
    b = RandomProjectionModule(device=DEV, decay_mode='lazy', **kw).to(DEV)
    b.random_projections[0].data.copy_(a.random_projections[0].data)
    b.mlp.load_state_dict(a.mlp.state_dict())
    t = 0.0
    cur = torch.cuda.current_stream(DEV)
    fs = a.feature_stream()
    for it in range(12):
        s = (1 + (rng.zipf(1.3, B) - 1) % (N - 1)).astype(np.int64)
        d = rng.integers(1, N, B).astype(np.int64)
        neg = rng.integers(1, N, B).astype(np.int64)
        ts = np.sort(t + rng.random(B) * 5000.0)
        t = float(ts[-1])
        if it % 3 == 2:                                      # device-resident ids
            s, d, neg, ts = (torch.from_numpy(x).to(DEV) for x in (s, d, neg, ts))
        with torch.no_grad():
            want_pos = b.get_pair_wise_feature(s, d)
            want_neg = b.get_pair_wise_feature(s, neg)
            a.update_prepare(s, d, ts, next_time=t)
            fs.wait_stream(cur)
            with torch.cuda.stream(fs):
                got_neg = a.get_pair_wise_feature(s, neg)
            got_pos = a.get_pair_wise_feature(s, d)
            cur.wait_stream(fs)
            got_neg.record_stream(cur)
        a.update(s, d, ts, next_time=t)
        b.update(s, d, ts, next_time=t)
        assert torch.equal(got_pos, want_pos), it
        assert torch.equal(got_neg, want_neg), it
    a.materialize()
    b.materialize()
    for i in range(1, L + 1):
        assert torch.equal(a.random_projections[i].data, b.random_projections[i].data), i
    assert a._h.stager2 is not None                          # the feature stream staged through its own ring
    a.check_errors()
    b.check_errors()
