"""CPU suite for the N>1 path: the routing plan and the one collective (all_to_all_single,
gloo backend, world_size 2 and 3) against the single-process oracle.  The rank-local kernels
are CUDA-only; here their role is played by a numpy executor that applies the plan's messages
with the oracle's arithmetic, so what is tested is exactly the host logic + the exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.walk_projection import WalkProjectionOracle, decay_factors, edge_weights
from tpnet_b200.sharded import (NativePlanner, exchange_blocks, make_plan, owner_of, rows_on_rank, update_messages)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_plan_is_consistent_across_ranks():
    rng = np.random.default_rng(0)
    N, B = 101, 400
    src = rng.integers(1, N, B); dst = rng.integers(1, N, B)
    tgt, oth, _ = update_messages(src, dst, np.zeros(B))
    for G in (1, 2, 3, 8):
        plans = [make_plan(tgt, oth, G, r, rows_on_rank(N, G, r)) for r in range(G)]
        assert sum(len(p.keep) for p in plans) == 2 * B            # every message has exactly one owner
        for r in range(G):
            p = plans[r]
            assert np.all(owner_of(tgt[p.keep], G) == r)
            assert np.all(np.diff(p.keep) > 0)                     # batch order preserved
            assert p.send_counts[r] == 0 and p.recv_counts[r] == 0
            for q in range(G):                                     # r's send list to q == q's receive list from r
                assert p.send_counts[q] == plans[q].recv_counts[r]
            # received slots are de-duplicated: one slot per distinct remote node
            remote = owner_of(oth[p.keep], G) != r
            assert p.num_recv == len(np.unique(oth[p.keep][remote]))
            assert np.all(p.second_rows[~remote] == oth[p.keep][~remote] // G)
            assert np.all(p.second_rows[remote] >= rows_on_rank(N, G, r))


def _same_plan(a, b):
    return (np.array_equal(a.keep, b.keep) and np.array_equal(a.first_rows, b.first_rows)
            and np.array_equal(a.second_rows, b.second_rows) and np.array_equal(a.send_rows, b.send_rows)
            and list(a.send_counts) == list(b.send_counts) and list(a.recv_counts) == list(b.recv_counts))


@pytest.mark.parametrize('G', [2, 3, 4, 8])
def test_native_planner_equals_the_numpy_plan(G):
    """csrc/tpn_plan.cu (what ShardedRandomProjection uses) vs make_plan, field for field: random and zipf
    batches, duplicates, self pairs, everything local / everything remote, empty input, reuse of one planner."""
    rng = np.random.default_rng(G)
    N = 5003
    planners = [NativePlanner(N, G, r) for r in range(G)]
    cases = []
    for B in (1, 7, 400, 20000):
        cases.append((rng.integers(0, N, B), rng.integers(0, N, B)))
        cases.append(((rng.zipf(1.3, B) - 1) % N, (rng.zipf(1.3, B) - 1) % N))
    a = rng.integers(0, N, 300)
    cases.append((a, a.copy()))                                         # self pairs: nothing crosses
    cases.append((np.full(500, G), np.arange(500) * G + 1 if G > 1 else np.arange(500)))   # one owner, all remote
    cases.append((np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)))
    for first, second in cases + cases[:3]:                             # again: the scratch state must be clean
        first, second = first.astype(np.int64), second.astype(np.int64)
        for r in range(G):
            n_local = rows_on_rank(N, G, r)
            want = make_plan(first, second, G, r, n_local)
            got = planners[r].plan(first, second, n_local)
            assert _same_plan(got, want), (G, r, len(first))
    with pytest.raises(IndexError):
        planners[0].plan(np.array([1, N], dtype=np.int64), np.array([2, 3], dtype=np.int64), rows_on_rank(N, G, 0))
    # ... and the failed call left nothing behind
    f, s2 = cases[1]
    assert _same_plan(planners[0].plan(f.astype(np.int64), s2.astype(np.int64), rows_on_rank(N, G, 0)),
                      make_plan(f.astype(np.int64), s2.astype(np.int64), G, 0, rows_on_rank(N, G, 0)))


def test_sharded_module_plans_with_the_native_builder():
    """ShardedRandomProjection.plan_update / plan_pairs (the calls bench.py and the public API make) on a stand-in
    object: world > 1 goes through the library's planner and yields the numpy plan; world == 1 stays in numpy."""
    from types import SimpleNamespace
    from tpnet_b200.sharded import ShardedRandomProjection as S
    rng = np.random.default_rng(9)
    N, B = 1001, 3000
    src, dst = rng.integers(1, N, B), rng.integers(1, N, B)
    t = np.sort(rng.random(B))
    for world in (1, 2, 4):
        for rank in range(world):
            me = SimpleNamespace(world=world, rank=rank, n_local=rows_on_rank(N, world, rank), global_node_num=N,
                                 _planner=None)
            me._make_plan = lambda a, b, me=me: S._make_plan(me, a, b)
            plan, tmsg = S.plan_update(me, src, dst, t)
            tgt, oth, tm = update_messages(src, dst, t)
            want = make_plan(tgt, oth, world, rank, me.n_local)
            assert _same_plan(plan, want) and np.array_equal(tmsg, tm[want.keep])
            assert (me._planner is None) == (world == 1)
            pp = S.plan_pairs(me, src, dst)
            assert _same_plan(pp, make_plan(src, dst, world, rank, me.n_local))
            for arr in (plan.keep, plan.first_rows, plan.second_rows, plan.send_rows):
                assert arr.dtype == np.int64 and arr.ndim == 1
                torch.from_numpy(np.ascontiguousarray(arr))          # what bench.py / the stager do with them


def _worker(rank, world, port, N, d, L, lam, batches, p0, result_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n_local = rows_on_rank(N, world, rank)
    ext = 4 * max(len(b[0]) for b in batches)
    W = (L + 1) * d
    state = np.zeros((n_local + ext, L + 1, d), dtype=np.float32)       # node-major, like the CUDA state
    state[:n_local, 0] = p0[rank::world]
    now = 0.0
    for src, dst, t in batches:
        tgt, oth, tm = update_messages(src, dst, t)
        plan = make_plan(tgt, oth, world, rank, n_local)
        send = torch.from_numpy(state[plan.send_rows].reshape(-1, W).copy())
        recv = torch.empty(plan.num_recv, W)
        exchange_blocks(send, plan.send_counts, recv, plan.recv_counts)
        state[n_local:n_local + plan.num_recv] = recv.numpy().reshape(-1, L + 1, d)
        # rank-local update with the oracle's arithmetic (the CUDA kernel's job on a GPU)
        t_last = t[-1]
        w = edge_weights(tm[plan.keep], t_last, lam)
        c = decay_factors(lam, t_last, now, L)
        for i in range(1, L + 1):
            state[:, i] = state[:, i] * c[i - 1]
        for i in range(L, 0, -1):
            msgs = state[plan.second_rows, i - 1] * w[:, None]
            for j in range(len(plan.keep)):
                state[plan.first_rows[j], i] += msgs[j]
        now = t_last
    np.save(os.path.join(result_dir, f'rank{rank}.npy'), state[:n_local])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_update_matches_oracle_with_gloo(world, tmp_path):
    rng = np.random.default_rng(world)
    N, d, L, lam = 61, 6, 3, 1e-3
    kw = dict(node_num=N, edge_num=500, dim_factor=1, num_layer=L, time_decay_weight=lam, use_matrix=False,
              beginning_time=0.0, not_scale=False, enforce_dim=d)
    o = WalkProjectionOracle(**kw)
    batches, t = [], 0.0
    for _ in range(5):
        B = 40
        s = 1 + (rng.zipf(1.4, B) - 1) % (N - 1)
        dd = 1 + (rng.zipf(1.4, B) - 1) % (N - 1)
        ts = np.sort(t + rng.random(B) * 200.0)
        t = ts[-1]
        batches.append((s.astype(np.int64), dd.astype(np.int64), ts))
        o.update(s, dd, ts)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, N, d, L, lam, batches, o.P[0].copy(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f'rank{r}.npy')
        for i in range(L + 1):
            assert np.array_equal(got[:, i], o.P[i][r::world]), f'rank {r} layer {i}'
