"""Node-sharded walk-projection state for graphs too large for one GPU.

The reference is single-device; this file is the multi-GPU extension `north_star` asks for.
Rows of node ``u`` live on rank ``u % G`` at local row ``u // G`` (modulo placement balances
power-law hubs).  The edge batch is replicated on every rank (24 B/edge).  Per call:

  1. every rank derives the SAME routing plan from the replicated batch (``ShardPlan``): which
     messages / pairs it owns (those whose target / first endpoint it owns, in batch order),
     which of its rows other ranks need, and which remote rows it needs — remote rows are
     de-duplicated per (owner, node), so a hub row crosses NVLink once per rank and batch;
  2. senders pack whole node blocks (rows 0..L, brought current) with ``tpn_gather_blocks``;
  3. ONE ``all_to_all_single`` (NCCL over NVLink/NVSwitch; gloo in the CPU tests) delivers the
     blocks straight into the extension rows that follow the local rows of the state buffer,
     so the kernels address local and received rows uniformly;
  4. the local kernels run: ``tpn_update_messages`` (pre-batch snapshot of the local targets +
     one all-layer walk launch; received rows are read in place; same per-row accumulation
     order as a single GPU, hence bit-identical results) or ``tpn_pairwise``.

Only the exchange is a collective; everything else is rank-local.  The plan is computed on
the host with numpy from the replicated batch (it can be precomputed for a resident batch).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .random_projection import RandomProjectionModule


def owner_of(ids: np.ndarray, world: int) -> np.ndarray:
    return ids % world


def local_row(ids: np.ndarray, world: int) -> np.ndarray:
    return ids // world


def rows_on_rank(node_num: int, world: int, rank: int) -> int:
    """Number of global ids u < node_num with u % world == rank."""
    return (node_num - rank + world - 1) // world if node_num > rank else 0


@dataclass
class ShardPlan:
    """Routing of one call, identical on every rank up to the rank-specific parts."""
    keep: np.ndarray            # indices (into the call's message / pair list) owned by this rank, in order
    first_rows: np.ndarray      # int64[M] local row of the target (update) / a endpoint (pairwise)
    second_rows: np.ndarray     # int64[M] row of the source / b endpoint: local row, or n_local + receive slot
    send_rows: np.ndarray       # int64[S] local rows to pack, grouped by destination rank (ascending), then node id
    send_counts: List[int]      # rows per destination rank
    recv_counts: List[int]      # rows per source rank (they land at n_local + offset, grouped by rank, then node id)

    @property
    def num_recv(self) -> int:
        return int(sum(self.recv_counts))


def make_plan(first: np.ndarray, second: np.ndarray, world: int, rank: int, n_local: int) -> ShardPlan:
    """Plan for messages/pairs (first[m], second[m]): work item m belongs to owner(first[m]) and
    needs the rows of second[m].  Pure function of the replicated inputs: rank r's send list to
    rank q is, by construction, rank q's receive list from rank r (same set, same order)."""
    first = np.asarray(first, dtype=np.int64)
    second = np.asarray(second, dtype=np.int64)
    if world == 1:                                                   # everything is local
        return ShardPlan(keep=np.arange(first.shape[0]), first_rows=first, second_rows=second,
                         send_rows=np.zeros(0, dtype=np.int64), send_counts=[0], recv_counts=[0])
    of, os_ = first % world, second % world
    big = np.int64(max(int(first.max(initial=0)), int(second.max(initial=0))) + 1)

    keep = np.nonzero(of == rank)[0]
    sec = second[keep]
    sec_owner = os_[keep]
    remote = sec_owner != rank
    second_rows = sec // world
    recv_counts = [0] * world
    if remote.any():
        key = sec_owner[remote] * big + sec[remote]                  # group by owner, then node id
        uniq, inverse = np.unique(key, return_inverse=True)
        second_rows = second_rows.copy()
        second_rows[remote] = n_local + inverse
        recv_counts = np.bincount(uniq // big, minlength=world).astype(np.int64).tolist()

    need = (os_ == rank) & (of != rank)                              # my rows that other ranks need
    send_counts = [0] * world
    send_rows = np.zeros(0, dtype=np.int64)
    if need.any():
        key = of[need] * big + second[need]                          # group by destination, then node id
        uniq = np.unique(key)
        send_rows = (uniq % big) // world
        send_counts = np.bincount(uniq // big, minlength=world).astype(np.int64).tolist()
    return ShardPlan(keep=keep, first_rows=first[keep] // world, second_rows=second_rows,
                     send_rows=send_rows.astype(np.int64), send_counts=[int(c) for c in send_counts],
                     recv_counts=[int(c) for c in recv_counts])


class NativePlanner:
    """`make_plan` computed by the library (csrc/tpn_plan.cu: two linear passes, no sort of the batch) —
    same plan, field for field (tests/test_sharded_cpu.py), 10x faster on 100k-edge batches."""

    def __init__(self, global_node_num: int, world: int, rank: int):
        self._lib = _lib.load()
        self.world, self.rank = int(world), int(rank)
        self.handle = ctypes.c_void_p()
        _lib.check(self._lib.tpn_planner_create(ctypes.byref(self.handle), int(global_node_num), self.world, self.rank),
                   'tpn_planner_create')

    def plan(self, first: np.ndarray, second: np.ndarray, n_local: int) -> ShardPlan:
        first = np.ascontiguousarray(first, dtype=np.int64)
        second = np.ascontiguousarray(second, dtype=np.int64)
        m = int(first.shape[0])
        if second.shape[0] != m:
            raise ValueError('first and second must have the same length')
        keep, rows1, rows2, send = (np.empty(max(m, 1), dtype=np.int64) for _ in range(4))
        sc, rc_ = np.zeros(self.world, dtype=np.int64), np.zeros(self.world, dtype=np.int64)
        nk, ns = ctypes.c_int64(0), ctypes.c_int64(0)
        rc = self._lib.tpn_plan(self.handle, first.ctypes.data, second.ctypes.data, m, int(n_local), keep.ctypes.data,
                                rows1.ctypes.data, rows2.ctypes.data, ctypes.byref(nk), send.ctypes.data,
                                ctypes.byref(ns), sc.ctypes.data, rc_.ctypes.data)
        if rc == _lib.TPN_ERR_INDEX:
            raise IndexError('node id out of range in a sharded batch')
        _lib.check(rc, 'tpn_plan')
        return ShardPlan(keep=keep[:nk.value], first_rows=rows1[:nk.value], second_rows=rows2[:nk.value],
                         send_rows=send[:ns.value], send_counts=[int(c) for c in sc], recv_counts=[int(c) for c in rc_])

    def __del__(self):
        try:
            if self.handle:
                self._lib.tpn_planner_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:  # interpreter shutdown
            pass


def update_messages(src: np.ndarray, dst: np.ndarray, t: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The 2B messages of one edge batch in the reference's accumulation order: all
    (target=src, source=dst) in batch order, then all (target=dst, source=src) — the two
    scatter_add_ calls of models/TPNet.py:93-96."""
    return np.concatenate([src, dst]), np.concatenate([dst, src]), np.concatenate([t, t])


def exchange_blocks(send: torch.Tensor, send_counts: Sequence[int], recv: torch.Tensor, recv_counts: Sequence[int],
                    group=None) -> None:
    """The one collective of the sharded path: rows of `send` ([S, W], grouped by destination
    rank) are delivered into `recv` ([R, W], grouped by source rank)."""
    if sum(send_counts) != send.shape[0] or sum(recv_counts) != recv.shape[0]:
        raise ValueError('split sizes do not match the buffers')
    dist.all_to_all_single(recv, send, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts),
                           group=group)


class ShardedRandomProjection(RandomProjectionModule):
    """Rank-local part of the node-sharded state.  Same constructor as the reference class plus
    the process group; ``node_num`` is the GLOBAL node count.  ``update`` and
    ``pair_wise_gram`` take the replicated global-id batch; ``pair_wise_gram`` returns the
    features of the pairs this rank owns together with their positions in the batch."""

    def __init__(self, node_num: int, edge_num: int, dim_factor: int, num_layer: int, time_decay_weight: float,
                 device: str, use_matrix: bool, beginning_time: np.float64, not_scale: bool, enforce_dim: int,
                 decay_mode: str = 'lazy', group=None, ext_rows: int = 1 << 16, p0: str = 'global',
                 state_device=None):
        if use_matrix:
            raise ValueError('use_matrix keeps N x N matrices: not meaningful for a sharded state')
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.global_node_num = int(node_num)
        self.n_local = rows_on_rank(self.global_node_num, self.world, self.rank)
        self.ext_rows = int(ext_rows)
        # the width rule uses the GLOBAL node count (TPNet.py:30-33)
        dim = enforce_dim if enforce_dim != -1 else min(int(math.log(edge_num * 2)) * dim_factor, node_num)
        self._p0_mode = p0
        if p0 == 'global':
            # draw the global P_0 exactly like the single-GPU module and keep this rank's rows:
            # same values as an unsharded run with the same seed (parity tests)
            full = torch.normal(0, 1 / math.sqrt(dim), (self.global_node_num, dim))
            mine = full[self.rank::self.world].clone()
            del full
        else:
            mine = None
        super().__init__(node_num=self.n_local + self.ext_rows, edge_num=edge_num, dim_factor=dim_factor,
                         num_layer=num_layer, time_decay_weight=time_decay_weight, device=device, use_matrix=False,
                         beginning_time=beginning_time, not_scale=not_scale, enforce_dim=dim, decay_mode=decay_mode,
                         init_p0=False, state_device=state_device)
        with torch.no_grad():
            if mine is not None:
                self.random_projections[0][:self.n_local].copy_(mine)
        self.exchanged_rows = 0          # rows received so far (bench accounting)
        self._send_buf: Optional[torch.Tensor] = None
        self._all_keep: Optional[np.ndarray] = None
        self._planner: Optional[NativePlanner] = None

    # ------------------------------------------------------------------ helpers
    def init_p0_on_device(self, seed: int) -> None:
        """P_0 ~ N(0, 1/sqrt(d)) drawn directly on this rank's GPU (large graphs: no host pass)."""
        g = torch.Generator(device=self._state.device).manual_seed(seed * 1000003 + self.rank)
        with torch.no_grad():
            p0 = self.random_projections[0]
            block = 1 << 20
            for lo in range(0, self.n_local, block):
                hi = min(lo + block, self.n_local)
                p0[lo:hi].copy_(torch.randn(hi - lo, self.dim, device=p0.device, generator=g) / math.sqrt(self.dim))

    def _ext_view(self, rows: int) -> torch.Tensor:
        if rows > self.ext_rows:
            raise RuntimeError(f'{rows} remote rows needed but the extension region holds {self.ext_rows}: '
                               f'construct ShardedRandomProjection with a larger ext_rows')
        return self._state.view(self.node_num, self.node_stride)[self.n_local:self.n_local + rows]

    def _exchange(self, plan: ShardPlan) -> None:
        """Pack -> all-to-all -> received blocks sit in the extension rows."""
        self._require_cuda()
        lib = _lib.load()
        dev = self._state.device
        S, R = int(plan.send_rows.shape[0]), plan.num_recv
        if self._send_buf is None or self._send_buf.shape[0] < max(S, 1):
            self._send_buf = torch.empty(max(S, 1024), self.node_stride, dtype=torch.float32, device=dev)
        send = self._send_buf[:S]
        if S:
            ptr = self._ids_to_device([plan.send_rows], ['id'])[0]
            rc = lib.tpn_gather_blocks(self._c_state(), ptr, S, send.data_ptr(), self._stream())
            if rc:
                _lib.check(rc, 'tpn_gather_blocks')
        recv = self._ext_view(R)
        if self.world > 1:
            exchange_blocks(send, plan.send_counts, recv, plan.recv_counts, self.group)
        if self.lazy and R:
            self._c_state()                                   # makes sure the stamps exist
            self._stamps[self.n_local:self.n_local + R].fill_(self._h.epoch)     # received rows are current
        self.exchanged_rows += R

    # ------------------------------------------------------------------ API
    def _make_plan(self, first: np.ndarray, second: np.ndarray) -> ShardPlan:
        if self.world == 1:
            return make_plan(first, second, 1, 0, self.n_local)
        if self._planner is None:
            self._planner = NativePlanner(self.global_node_num, self.world, self.rank)
        return self._planner.plan(first, second, self.n_local)

    def plan_update(self, src: np.ndarray, dst: np.ndarray, t: np.ndarray):
        tgt, oth, tm = update_messages(np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64),
                                       np.asarray(t, dtype=np.float64))
        plan = self._make_plan(tgt, oth)
        return plan, np.ascontiguousarray(tm[plan.keep])

    def update(self, src_node_ids, dst_node_ids, node_interact_times, next_time=None, plan=None):
        """TPNet.py:67-99 on the sharded state.  All ranks must call it with the same batch."""
        if self.world == 1 and plan is None:
            # one shard = the whole graph: the plain edge-batch path (no message list, no exchange)
            return RandomProjectionModule.update(self, src_node_ids, dst_node_ids, node_interact_times, next_time)
        dev = self._require_cuda()
        lib = _lib.load()
        t = np.asarray(node_interact_times, dtype=np.float64)
        if len(t) == 0:
            raise IndexError('index -1 is out of bounds for axis 0 with size 0')
        if plan is None:
            src = np.asarray(src_node_ids, dtype=np.int64)
            dst = np.asarray(dst_node_ids, dtype=np.int64)
            if src.min() < 0 or dst.min() < 0 or src.max() >= self.global_node_num or \
                    dst.max() >= self.global_node_num:
                raise IndexError(f'index out of range for node_num {self.global_node_num}')
            plan = self.plan_update(src, dst, t)
        plan, t_msg = plan
        next_time = float(t[-1]) if next_time is None else float(next_time)
        h = self._h
        lam = self.time_decay_weight
        base = np.exp(-lam * (np.float64(next_time) - np.float64(h.now)))
        factors = (ctypes.c_float * self.num_layer)(*[float(np.float32(np.power(base, i)))
                                                      for i in range(1, self.num_layer + 1)])
        self._exchange(plan)                                  # pre-batch rows, before the clock moves
        M = int(plan.first_rows.shape[0])
        if M:
            ptrs = self._ids_to_device([plan.first_rows, plan.second_rows, t_msg], ['id', 'id', 'time'],
                                       wrap_negative=False)
            st = self._c_state()
            need = lib.tpn_update_workspace_bytes(st, (M + 1) // 2)
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=dev)
            if self._err is None:
                self._err = torch.zeros(1, dtype=torch.int32, device=dev)
            args = (ptrs[0], ptrs[1], ptrs[2], M, self.n_local, next_time, float(np.float32(-lam)), factors,
                    self._ws.data_ptr(), self._ws.numel(), self._err.data_ptr(), self._stream())
            rc = lib.tpn_update_messages(st, *args)
            if rc == _lib.TPN_ERR_LOG_FULL:
                self._restart_log()
                st = self._c_state()
                rc = lib.tpn_update_messages(st, *args)
            if rc:
                _lib.check(rc, 'tpn_update_messages')
            h.epoch = int(h.st.epoch)
            h.launches += 1
        else:
            # no local message: the clock (and the lazy epoch) must still advance identically
            self._advance_clock_only(factors)
        h.now = next_time
        self.now_time.data.fill_(h.now)

    def _advance_clock_only(self, factors) -> None:
        """Decay without any local message (keeps epochs aligned across ranks)."""
        if all(f == 1.0 for f in factors):
            return
        if self.lazy:
            self._c_state()
            if self._h.epoch + 1 >= self._decay_log.shape[0]:
                self._restart_log()
            if self._h.st.cum_floor * min(factors) < 1e-200:
                self._restart_log()
            self._h.epoch += 1
            f = torch.tensor([float(x) for x in factors], dtype=torch.float64, device=self._decay_log.device)
            self._decay_log[self._h.epoch] = self._decay_log[self._h.epoch - 1] * f      # cumulative products
            self._h.st.cum_floor *= float(min(factors))
        else:
            with torch.no_grad():
                for i in range(1, self.num_layer + 1):
                    self.random_projections[i].mul_(float(factors[i - 1]))

    def plan_pairs(self, a_ids: np.ndarray, b_ids: np.ndarray) -> ShardPlan:
        return self._make_plan(np.asarray(a_ids, dtype=np.int64), np.asarray(b_ids, dtype=np.int64))

    def pair_wise_gram(self, src_node_ids, dst_node_ids, plan: Optional[ShardPlan] = None):
        """Features (input of self.mlp, TPNet.py:119-128) of the pairs whose first endpoint this
        rank owns.  Returns (positions in the batch, float32 [m, (2L+2)^2])."""
        if self.world == 1 and plan is None:
            n = int(len(src_node_ids))
            if self._all_keep is None or self._all_keep.shape[0] != n:
                self._all_keep = np.arange(n)
            return self._all_keep, RandomProjectionModule.pair_wise_gram(self, src_node_ids, dst_node_ids)
        dev = self._require_cuda()
        lib = _lib.load()
        if plan is None:
            plan = self.plan_pairs(src_node_ids, dst_node_ids)
        self._exchange(plan)
        m = int(plan.first_rows.shape[0])
        out = torch.empty(m, self.pair_wise_feature_dim, dtype=torch.float32, device=dev)
        if m:
            ptrs = self._ids_to_device([plan.first_rows, plan.second_rows], ['id', 'id'])
            rc = lib.tpn_pairwise(self._c_state(), ptrs[0], ptrs[1], m, 0 if self.not_scale else 1, out.data_ptr(),
                                  self._stream())
            if rc:
                _lib.check(rc, 'tpn_pairwise')
            self._h.launches += 1
        return plan.keep, out

    def get_pair_wise_feature(self, src_node_ids, dst_node_ids):
        keep, feat = self.pair_wise_gram(src_node_ids, dst_node_ids)
        return keep, self._head(feat)

    def gather_global(self) -> List[torch.Tensor]:
        """All-gathers the L+1 global [N, d] matrices (tests / small graphs only)."""
        self.materialize()
        out = []
        for i in range(self.num_layer + 1):
            mine = self.random_projections[i].data[:self.n_local].contiguous()
            full = torch.zeros(self.global_node_num, self.dim, dtype=torch.float32, device=mine.device)
            if self.world == 1:
                full.copy_(mine)
            else:
                pad = rows_on_rank(self.global_node_num, self.world, 0)
                buf = torch.zeros(pad, self.dim, dtype=torch.float32, device=mine.device)
                buf[:self.n_local].copy_(mine)
                parts = [torch.empty_like(buf) for _ in range(self.world)]
                dist.all_gather(parts, buf, group=self.group)
                for r in range(self.world):
                    n_r = rows_on_rank(self.global_node_num, self.world, r)
                    full[r::self.world] = parts[r][:n_r]
            out.append(full)
        return out
