"""CPU suite: the oracle against the committed reference fixtures and the
brute-force walk enumeration.  No GPU, no /root/reference."""
import numpy as np
import pytest

from golden_util import CASES, load_case, oracle_kwargs
from oracle.walk_bruteforce import random_temporal_graph, sum_walk_matrices
from oracle.walk_projection import WalkProjectionOracle, decay_factors, edge_weights, projection_dim


@pytest.mark.parametrize('name', CASES)
def test_oracle_bit_exact_against_reference_fixture(name):
    z, cfg, batches = load_case(name)
    kw = oracle_kwargs(cfg)
    o = WalkProjectionOracle(p0=None if kw['use_matrix'] else z['p0'], **kw)
    assert o.dim == int(z['dim'])
    L = kw['num_layer']
    for b, (s, d, t, w) in enumerate(batches):
        o.update(s, d, t, weights=w)                 # the reference's own torch-exp weights
        if b == 0:
            for i in range(1, L + 1):
                assert np.array_equal(o.P[i], z[f'after0_P{i}'])
        if 'backup_now' in z.files and False:
            pass
    for i in range(1, L + 1):
        assert np.array_equal(o.P[i], z[f'final_P{i}']), f'layer {i}'
    assert o.now_time == z['final_now']
    assert np.array_equal(o.P[0], z['p0']), 'P_0 must never be written by update'


@pytest.mark.parametrize('name', CASES)
def test_oracle_own_weights_close_to_reference(name):
    z, cfg, batches = load_case(name)
    kw = oracle_kwargs(cfg)
    o = WalkProjectionOracle(p0=None if kw['use_matrix'] else z['p0'], **kw)
    for s, d, t, w in batches:
        w_own = edge_weights(t, t[-1], kw['time_decay_weight'])
        # torch's exp is 1-ulp: never more than one ulp away from the correctly rounded value
        assert np.all(np.abs(w_own.astype(np.float64) - w) <= np.spacing(np.maximum(w_own, w)))
        o.update(s, d, t)
    for i in range(1, kw['num_layer'] + 1):
        np.testing.assert_allclose(o.P[i], z[f'final_P{i}'], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize('name', CASES)
def test_oracle_pairwise_and_gather(name):
    z, cfg, batches = load_case(name)
    kw = oracle_kwargs(cfg)
    o = WalkProjectionOracle(p0=None if kw['use_matrix'] else z['p0'], **kw)
    for s, d, t, w in batches:
        o.update(s, d, t, weights=w)
    a, b, ref = z['pair_a'], z['pair_b'], z['pair_feat']
    got = o.pair_wise_gram(a, b)
    assert got.shape == ref.shape == (len(a), (2 * kw['num_layer'] + 2) ** 2)
    if kw['not_scale']:
        scale = o.pair_norm_bound(a, b)
        assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 2e-6 * scale)
    else:
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-6)
        assert np.all(ref >= 0)                       # clamp + log(x+1) is non-negative
    rows = o.get_random_projections(a)
    assert len(rows) == kw['num_layer'] + 1 and rows[0].shape == (len(a), o.dim)


@pytest.mark.parametrize('name', CASES)
def test_oracle_neighbor_pairwise_matches_reference_fixture(name):
    """TPNet.py:313-324 (index lists + re-split) restated in the oracle vs the reference's output."""
    z, cfg, batches = load_case(name)
    kw = oracle_kwargs(cfg)
    o = WalkProjectionOracle(p0=None if kw['use_matrix'] else z['p0'], **kw)
    for s, d, t, w in batches:
        o.update(s, d, t, weights=w)
    nbr, src, dst, ref = z['nbr'], z['nbr_src'], z['nbr_dst'], z['nbr_feat']
    m, k = nbr.shape
    F = (2 * kw['num_layer'] + 2) ** 2
    got = o.neighbor_pair_wise_gram(nbr, src, dst)
    assert got.shape == ref.shape == (m, k, 2 * F)
    a, b = o.neighbor_pair_lists(nbr, src, dst)
    assert len(a) == len(b) == 2 * m * k
    if kw['not_scale']:
        scale = o.pair_norm_bound(a, b)
        scale = np.concatenate([scale[:m * k], scale[m * k:]], axis=1).reshape(m, k, -1)
        assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 2e-6 * scale)
    else:
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-6)
    # block [n, k, 0] is the pair (nbr[n,k], src[n]); block [n, k, 1] the pair (nbr[n,k], dst[n])
    one = o.pair_wise_gram(nbr[4], np.repeat(dst[4], k))
    assert np.array_equal(got[4, :, F:], one)


@pytest.mark.parametrize('name', ['wiki_tiny', 'flights_tiny'])
def test_oracle_backup_reload(name):
    z, cfg, batches = load_case(name)
    kw = oracle_kwargs(cfg)
    L = kw['num_layer']
    o = WalkProjectionOracle(p0=z['p0'], **kw)
    nb = len(batches)
    saved = None
    for b, (s, d, t, w) in enumerate(batches):
        o.update(s, d, t, weights=w)
        if f'backup_P1' in z.files and saved is None:
            # find the batch the fixture backed up at by matching the clock
            if o.now_time == z['backup_now']:
                ok = all(np.array_equal(o.P[i], z[f'backup_P{i}']) for i in range(1, L + 1))
                if ok:
                    saved = o.backup()
    assert saved is not None
    p0_before = o.P[0].copy()
    o.reload(saved)
    assert o.now_time == z['backup_now']
    s, d, t, _ = batches[-1]
    o.update(s, d, t, weights=z['after_reload_w'])
    for i in range(1, L + 1):
        assert np.array_equal(o.P[i], z[f'after_reload_P{i}'])
    assert np.array_equal(o.P[0], p0_before)


def test_layers_consume_pre_batch_state():
    """Layer independence within a batch (reference processes layers top-down,
    TPNet.py:90): after the first batch P_2.. are still all zero."""
    z, cfg, batches = load_case('reddit_tiny')
    assert np.any(z['after0_P1'] != 0)
    assert not np.any(z['after0_P2']) and not np.any(z['after0_P3'])


def test_clock_is_last_element_not_max():
    kw = dict(node_num=6, edge_num=10, dim_factor=1, num_layer=1, time_decay_weight=0.1, use_matrix=False,
              beginning_time=0.0, not_scale=True, enforce_dim=4)
    o = WalkProjectionOracle(**kw)
    o.update(np.array([1, 2]), np.array([3, 4]), np.array([9.0, 5.0]))
    assert o.now_time == 5.0


def test_empty_batch_raises_like_reference():
    kw = dict(node_num=6, edge_num=10, dim_factor=1, num_layer=1, time_decay_weight=0.1, use_matrix=False,
              beginning_time=0.0, not_scale=True, enforce_dim=4)
    o = WalkProjectionOracle(**kw)
    e = np.array([], dtype=np.int64)
    with pytest.raises(IndexError):
        o.update(e, e, np.array([], dtype=np.float64))
    with pytest.raises(IndexError):
        o.update(np.array([1]), np.array([6]), np.array([1.0]))


def test_dim_rule():
    assert projection_dim(9228, 157475, 10, -1, False) == 120      # Wikipedia shape
    assert projection_dim(10985, 672448, 10, -1, False) == 140     # Reddit shape
    assert projection_dim(13170, 1927146, 10, -1, False) == 150    # Flights shape
    assert projection_dim(50, 157475, 10, -1, False) == 50         # capped by node_num
    assert projection_dim(50, 10, 10, 64, False) == 64
    assert projection_dim(50, 10, 10, 64, True) == 50


def test_decay_factor_rounding():
    c = decay_factors(1e-6, 3400.0, 0.0, 3)
    base = np.exp(-1e-6 * 3400.0)
    assert c.dtype == np.float32 and c[2] == np.float32(base ** 3)
    assert np.all(decay_factors(1e-6, 7.0, 7.0, 3) == 1.0)


def test_bruteforce_kat_fresh_graph():
    """The notebook's known-answer test on a freshly drawn graph (cell 10)."""
    rng = np.random.default_rng(7)
    N, E, L, lam, B = 16, 60, 3, 1e-4, 5
    s, d, _ = random_temporal_graph(N, E, rng)
    t = np.repeat(np.arange(1, E // B + 1), B).astype(np.float64)
    o = WalkProjectionOracle(node_num=N, edge_num=E, dim_factor=1, num_layer=L, time_decay_weight=lam,
                             use_matrix=True, beginning_time=0.0, not_scale=True, enforce_dim=-1)
    for i in range(0, E, B):
        o.update(s[i:i + B], d[i:i + B], t[i:i + B])
    brute = sum_walk_matrices(s, d, t, L, lam, N)
    for j in range(L + 1):
        shifted = o.P[j].astype(np.float64) * np.power(np.exp(-lam), j)
        np.testing.assert_allclose(shifted, brute[j], rtol=1e-5, atol=1e-5)


def test_bruteforce_fixture():
    z = np.load(__import__('os').path.join(__import__('golden_util').GOLDEN_DIR, 'matrix_kat_brute.npz'))
    brute = sum_walk_matrices(z['src'], z['dst'], z['t'], 3, 1e-4, 24)
    for j in range(4):
        np.testing.assert_allclose(brute[j], z[f'A{j}'], rtol=1e-12)


def test_add_at_equals_loop():
    """The large-batch path of the oracle (np.add.at) is the same sequential sum."""
    rng = np.random.default_rng(3)
    kw = dict(node_num=50, edge_num=4000, dim_factor=2, num_layer=3, time_decay_weight=1e-3, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=-1)
    a = WalkProjectionOracle(**kw)
    b = WalkProjectionOracle(p0=a.P[0], **kw)
    b.LOOP_MAX = 0
    t = 0.0
    for _ in range(4):
        B = 300
        s = 1 + (rng.zipf(1.4, B) - 1) % 49
        d = 1 + (rng.zipf(1.4, B) - 1) % 49
        ts = np.sort(t + rng.random(B) * 100)
        t = ts[-1]
        a.update(s, d, ts)
        b.update(s, d, ts)
    for i in range(4):
        assert np.array_equal(a.P[i], b.P[i])


# --------------------------------------------------------------------------- `recent` neighbour sampler
def _sampler_case():
    import os
    from golden_util import GOLDEN_DIR
    return np.load(os.path.join(GOLDEN_DIR, 'sampler_tiny.npz'))


def test_sampler_oracle_matches_reference_fixture():
    """oracle/neighbor_sampler.py vs the outputs of the reference's NeighborSampler ('recent', utils/utils.py:160-224)
    stored by tests/golden/make_golden.py: equal timestamps, a self loop, a repeated edge, node 0, K = 1, 5, 20."""
    from oracle.neighbor_sampler import RecentNeighborOracle
    z = _sampler_case()
    o = RecentNeighborOracle(z['src'], z['dst'], z['eid'], z['t'], int(z['num_nodes']))
    for k in range(3):
        got = o.get_historical_neighbors(z[f'q{k}_nodes'], z[f'q{k}_times'], int(z[f'q{k}_K']))
        for x, name in zip(got, ('nbr', 'eid', 't')):
            assert np.array_equal(x, z[f'q{k}_{name}']) and x.dtype == z[f'q{k}_{name}'].dtype
        assert not got[0][z[f'q{k}_nodes'] == 0].any()                  # the padding node has no history


def _kernel_arithmetic(offsets, nbr, eid, times, num_nodes, q_nodes, q_times, K):
    """The index arithmetic of csrc/tpn_sampler.cu (one query per warp), line for line, on the host."""
    n = len(q_nodes)
    out_n, out_e, out_t = np.zeros((n, K), np.int64), np.zeros((n, K), np.int64), np.zeros((n, K), np.float64)
    for q in range(n):
        node, t = int(q_nodes[q]), float(q_times[q])
        begin = end = 0
        if 0 <= node < num_nodes:
            begin, end = int(offsets[node]), int(offsets[node + 1])
        a, b = begin, end
        while a < b:
            mid = a + ((b - a) >> 1)
            if times[mid] < t:
                a = mid + 1
            else:
                b = mid
        take = min(a - begin, K)
        first, pad = a - take, K - take
        for j in range(K):
            if j >= pad:
                s = first + (j - pad)
                out_n[q, j], out_e[q, j], out_t[q, j] = nbr[s], eid[s], times[s]
    return out_n, out_e, out_t


def test_sampler_host_logic_and_kernel_arithmetic():
    """The product's CSR builder (host logic, tpnet_b200/neighbor_sampler.py) equals the oracle's adjacency, and
    the kernel's search / padding arithmetic over it reproduces the reference outputs."""
    from oracle.neighbor_sampler import RecentNeighborOracle
    from tpnet_b200.neighbor_sampler import RecentNeighborSampler, build_recent_csr
    z = _sampler_case()
    N = int(z['num_nodes'])
    off, nbr, eid, tt = build_recent_csr(z['src'], z['dst'], z['eid'], z['t'], N)
    o = RecentNeighborOracle(z['src'], z['dst'], z['eid'], z['t'], N)
    assert np.array_equal(off, o.offsets) and np.array_equal(nbr, o.nbr) and np.array_equal(eid, o.eid)
    assert np.array_equal(tt, o.times) and off[0] == 0 and off[1] == 0 and off[-1] == 2 * len(z['src'])
    for k in range(3):
        got = _kernel_arithmetic(off, nbr, eid, tt, N, z[f'q{k}_nodes'], z[f'q{k}_times'], int(z[f'q{k}_K']))
        for x, name in zip(got, ('nbr', 'eid', 't')):
            assert np.array_equal(x, z[f'q{k}_{name}'])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        RecentNeighborSampler(z['src'], z['dst'], z['eid'], z['t'], device='cpu')
