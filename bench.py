#!/usr/bin/env python
"""bench.py — temporal edges/sec through the walk-projection hot path (update + pair-wise
encode) on synthetic graphs of the BASELINE.json shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload powerlaw|reddit|wikipedia|flights]

Primary workload (all N): BASELINE.json configs[3], the one its metric ("... at 1/2/4/8 B200;
% HBM roofline") is quoted on — a power-law temporal graph with 10M nodes, d=210, L=3, lazy
decay, batches of 100,000 edges, p=2 pair-encodes per edge (the decoder shape of
models/modules.py:112: (src,dst) and (src,neg)).  It fits one GPU (34.6 GB state) and is
HBM-bound, so the roofline fraction is a real DRAM figure.  N>1: the state is sharded by node
id (tpnet_b200/sharded.py); weak scaling by default — every GPU brings 100,000 edges per step, the
job's batch is 100,000 x N (`--scaling strong` splits one 100,000-edge batch instead).
At N=1 the line also carries `also.reddit`: BASELINE.json configs[1] (Reddit-shaped, B=200,
K=20, p=162 — one TPNet training batch, latency-bound, state L2-resident).

One JSON line on stdout (driver contract).  `value`: inputs resident in HBM, CUDA-event time.
`e2e`: the public numpy API with pinned H2D of the step's inputs, `self.mlp`, and a scalar
result read back per step.  `--impl reference`: the torch-CPU port of the reference
(oracle/cpu_port.py) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tpnet_b200.synth import (SHAPES, RecentNeighbors, edge_stream, tpnet_neighbor_batch,  # noqa: E402
                               tpnet_pair_lists)

BATCH = 200                 # TPNet batch (configs[0..2])
NUM_NEIGHBORS = 20
WARM_BATCHES = 400          # untimed: populates P_1..P_L and the neighbour table
PL_BATCH = 100_000          # power-law batch (configs[3])
PL_WARM = 12
METRIC = 'temporal edges/sec (update+pairwise encode)'
REF_THREADS = 3             # the reference pins torch to 3 intra-op threads (train_link_prediction.py:124)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=None)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='powerlaw', choices=['powerlaw', 'reddit', 'wikipedia', 'flights'])
    ap.add_argument('--decay-mode', default='auto', choices=['auto', 'eager', 'lazy'])
    ap.add_argument('--no-flush', action='store_true', help='small shapes: keep L2 warm between steps')
    ap.add_argument('--no-graphs', action='store_true', help='power-law, 1 GPU: launch the steps eagerly instead of as CUDA graphs')
    ap.add_argument('--no-also', action='store_true', help='N=1: skip the secondary Reddit-shaped measurement')
    ap.add_argument('--cpu-sample-steps', type=int, default=None)
    ap.add_argument('--cpu-threads', type=int, default=None, help='--impl reference: host threads (default: all cores)')
    ap.add_argument('--warm-batches', type=int, default=None, help='untimed batches that fill the state')
    ap.add_argument('--pl-nodes', type=int, default=None, help='override the power-law node count (debug)')
    ap.add_argument('--pl-batch', type=int, default=PL_BATCH, help='power-law: edges per step and GPU')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='N > 1: weak = pl_batch edges per GPU and step (job batch pl_batch x N), strong = pl_batch in total')
    return ap.parse_args()


def peak_gbs():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (burst copy)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def algorithmic_bytes(shape):
    L, d = shape.num_layer, shape.dim
    per_edge = 24 * L * d + 24
    per_pair = 2 * (L + 1) * d * 4 + (2 * L + 2) ** 2 * 4 + 16
    return per_edge, per_pair


def traffic_note(key):
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(p):
        return json.load(open(p)).get(key)
    return None


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                                          '-lms', '50', '-i', str(index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [r for ts, r in self.rows if t0 - 0.15 <= ts <= t1 + 0.15] or [r for _, r in self.rows[-3:]]
        for r in rows:
            p = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


# ============================================================================= Reddit-shaped (TPNet batch) workload
def make_steps(shape, n_steps, seed, warm=WARM_BATCHES):
    """Returns (warm_batches, steps).  Each step holds the numpy inputs of one batch."""
    rng = np.random.default_rng(seed + 17)
    nbr = RecentNeighbors(shape.node_num, NUM_NEIGHBORS)
    stream = edge_stream(shape, BATCH, warm + n_steps, seed=seed)
    warm_batches, steps = [], []
    lo, hi = (shape.num_src + 1, shape.num_src + shape.num_dst + 1) if shape.num_dst else (1, shape.num_src + 1)
    for i, (s, d, t) in enumerate(stream):
        if i < warm:
            warm_batches.append((s, d, t))
        else:
            neg = rng.integers(lo, hi, BATCH).astype(np.int64)          # random negative sampling
            pa, pb = tpnet_pair_lists(nbr, s, d)
            na, nb = tpnet_pair_lists(nbr, s, neg)
            steps.append(dict(src=s, dst=d, t=t, neg=neg, enc_pos=(pa, pb), enc_neg=(na, nb),
                              nbr_pos=tpnet_neighbor_batch(nbr, s, d), nbr_neg=tpnet_neighbor_batch(nbr, s, neg)))
        nbr.insert(s, d)
    return warm_batches, steps


def pairs_per_step():
    return 8 * BATCH * NUM_NEIGHBORS + 2 * BATCH


def cpu_port_tpnet(shape, warm_batches, steps, n_warm, n_timed, threads):
    """Times the torch-CPU port of the reference on the host cores: same step shape."""
    from oracle.cpu_port import CpuWalkProjection          # baseline leg only (see oracle/__init__.py)
    torch.set_num_threads(threads)
    ref = CpuWalkProjection(shape.node_num, shape.dim, shape.num_layer, shape.time_decay_weight,
                            float(warm_batches[0][2][0]), not_scale=False, with_mlp=True, seed=0)
    for s, d, t in warm_batches[-40:]:                      # a short warm stream is enough to fill the rows
        ref.update(s, d, t)

    def one(st):
        with torch.no_grad():
            ref.pair_wise(*st['enc_pos'])
            ref.pair_wise(*st['enc_neg'])
            pos = ref.pair_wise(st['src'], st['dst'])
            neg = ref.pair_wise(st['src'], st['neg'])
            ref.update(st['src'], st['dst'], st['t'])
            res = pos.sum() - neg.sum()
        return float(res)

    for st in steps[:n_warm]:
        one(st)
    t0 = time.perf_counter()
    for st in steps[n_warm:n_warm + n_timed]:
        one(st)
    return (time.perf_counter() - t0) / n_timed


def build_module(shape, device, decay_mode, t0):
    from tpnet_b200 import RandomProjectionModule
    torch.manual_seed(0)
    m = RandomProjectionModule(node_num=shape.node_num, edge_num=shape.edge_num, dim_factor=shape.dim_factor,
                               num_layer=shape.num_layer, time_decay_weight=shape.time_decay_weight,
                               device=str(device), use_matrix=False, beginning_time=np.float64(t0), not_scale=False,
                               enforce_dim=-1, decay_mode=decay_mode)
    return m.to(device)


def to_dev(st, device):
    g = lambda a, dt: torch.from_numpy(a).to(device=device, dtype=dt)     # noqa: E731
    return dict(src=g(st['src'], torch.int64), dst=g(st['dst'], torch.int64), t=g(st['t'], torch.float64),
                neg=g(st['neg'], torch.int64), t_last=float(st['t'][-1]),
                nbr_pos=tuple(g(a, torch.int64) for a in st['nbr_pos']),
                nbr_neg=tuple(g(a, torch.int64) for a in st['nbr_neg']))


def resident_step(m, ds):
    """One step with device-resident inputs: kernels only (no head, no H2D).  The feature
    tensors are dropped at once: inside a capture their memory returns to the graph pool."""
    m.neighbor_pair_wise_gram(*ds['nbr_pos'])      # encoder: [2B, K, 2, F] = 4BK pair blocks (TPNet.py:313-324)
    m.neighbor_pair_wise_gram(*ds['nbr_neg'])
    m.pair_wise_gram(ds['src'], ds['dst'])
    m.pair_wise_gram(ds['src'], ds['neg'])
    m.update(ds['src'], ds['dst'], ds['t'], next_time=ds['t_last'])


def api_step(m, st):
    """One step through the public numpy API, head included, scalar result read back."""
    with torch.no_grad():
        m.get_neighbor_pair_wise_feature(*st['nbr_pos'])       # encoder features (feed the Mixer in TPNet)
        m.get_neighbor_pair_wise_feature(*st['nbr_neg'])
        pos = m.get_pair_wise_feature(st['src'], st['dst'])    # decoder features -> the step's scalar result
        neg = m.get_pair_wise_feature(st['src'], st['neg'])
        m.update(st['src'], st['dst'], st['t'])
        res = pos.sum() - neg.sum()
    return float(res.item())                                   # D2H read of the step's result


def run_tpnet_shape(args, shape, device, K, W, with_cpu=True):
    """Single-GPU TPNet-batch workload (Reddit / Wikipedia / Flights shapes)."""
    per_edge_B, per_pair_B = algorithmic_bytes(shape)
    pps = pairs_per_step()
    warm_n = args.warm_batches if args.warm_batches is not None else WARM_BATCHES
    warm_batches, steps = make_steps(shape, 2 * (K + W), seed=0, warm=warm_n)
    m = build_module(shape, device, args.decay_mode, warm_batches[0][2][0])
    for s, d, t in warm_batches:
        m.update(s, d, t)
    torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    dsteps = [to_dev(st, device) for st in steps[:K + W]]
    graphs = []
    side = torch.cuda.Stream(device)
    pool = torch.cuda.graph_pool_handle()       # graphs replay in capture order, so they can share memory
    with torch.cuda.stream(side):
        for ds in dsteps:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=side):
                resident_step(m, ds)
            graphs.append(g)
    torch.cuda.synchronize()
    for g in graphs[:W]:
        g.replay()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    wall0 = time.perf_counter()
    for k in range(K):
        if not args.no_flush:
            flush.zero_()                                   # evict the 126 MB L2 between steps (untimed)
        ev[k][0].record()
        graphs[W + k].replay()
        ev[k][1].record()
    torch.cuda.synchronize()
    wall1 = time.perf_counter()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    m.check_errors()

    # dominant kernel (the 4BK-pair encoder launch): each launch alone in a CUDA graph so that no
    # host launch latency sits between the two events; L2 flushed before each
    n_pair = min(K, 100)
    pair_graphs = []
    with torch.cuda.stream(side):
        for k in range(n_pair):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=side):
                m.neighbor_pair_wise_gram(*dsteps[W + k]['nbr_pos'])
            pair_graphs.append(g)
    torch.cuda.synchronize()
    pev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_pair)]
    for k in range(n_pair):
        flush.zero_()
        pev[k][0].record()
        pair_graphs[k].replay()
        pev[k][1].record()
    torch.cuda.synchronize()
    pair_ms = float(np.mean([a.elapsed_time(b) for a, b in pev]))
    pairs_per_launch = 2 * dsteps[0]['nbr_pos'][0].numel()        # two pair blocks per (row, neighbour)

    api = steps[K + W:]
    for st in api[:W]:
        api_step(m, st)
    torch.cuda.synchronize()
    e0 = time.perf_counter()
    for st in api[W:W + K]:
        api_step(m, st)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - e0) * 1e3
    h2d = 2 * ((2 * BATCH * NUM_NEIGHBORS + 4 * BATCH) * 8) + 2 * (2 * BATCH * 8) + 3 * BATCH * 8

    peak, peak_src = peak_gbs()
    achieved = pairs_per_launch * per_pair_B / (pair_ms * 1e-3) / 1e9
    step_bytes = BATCH * per_edge_B + pps * per_pair_B
    kernels_per_step = 4 + 3                                  # 4 pair-wise + prep(+sweep) + snapshot + walk
    out = {
        'value': BATCH * K / (dev_ms * 1e-3), 'unit': 'edges/s', 'steps': K, 'warmup': W, 'ms_per_step': dev_ms / K,
        'config': {'workload': (f'{shape.name}-shaped synthetic graph ({shape.num_nodes} nodes, {shape.num_edges} '
                                f'edges), d={shape.dim}, L={shape.num_layer}, batch {BATCH}, K={NUM_NEIGHBORS}, '
                                f'{pps // BATCH} pair-encodes per edge, random negatives'),
                   'batch': BATCH, 'num_neighbors': NUM_NEIGHBORS, 'pairs_per_step': pps, 'dim': shape.dim,
                   'num_layer': shape.num_layer, 'node_num': shape.node_num, 'parallelism': 'single GPU',
                   'decay_mode': 'lazy' if m.lazy else 'eager',
                   'l2': 'warm (no flush)' if args.no_flush else 'flushed between steps (256 MiB memset, untimed)',
                   'timing': 'CUDA events per step, one CUDA graph per step',
                   'algorithmic_bytes_per_step': step_bytes,
                   'wall_ms_per_step_incl_flush': (wall1 - wall0) * 1e3 / K},
        'pairs_per_s': pps * K / (dev_ms * 1e-3),
        'algorithmic_GBps_step': step_bytes * K / (dev_ms * 1e-3) / 1e9,
        'roofline': {'bound': 'hbm', 'kernel': 'tpn::pairwise_nbr_kernel (encoder launch: 2B rows x K neighbours = 4BK pair blocks)',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic_note('reddit_pairwise_dram_bytes_per_launch'), 'peak_source': peak_src,
                     'launch_us': pair_ms * 1e3, 'pairs_per_launch': pairs_per_launch,
                     'note': 'state (%.1f MB) is L2-resident: algorithmic bytes are served from L2 after first '
                             'touch, so this is a fraction of the HBM copy peak reached out of L2, not DRAM use' %
                             (shape.node_num * (shape.num_layer + 1) * m.row_stride * 4 / 1e6)},
        'e2e': {'value': BATCH * K / (e2e_ms * 1e-3), 'unit': 'edges/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': 4, 'ms_per_step': e2e_ms / K,
                'path': 'RandomProjectionModule.get_pair_wise_feature/update with numpy ids, self.mlp included'},
        'gpu_launches': kernels_per_step * K,
    }
    if with_cpu:
        threads = os.cpu_count() or 1
        n_cpu = args.cpu_sample_steps or 30
        sec = cpu_port_tpnet(shape, warm_batches, steps, 2, n_cpu, threads)
        sec3 = cpu_port_tpnet(shape, warm_batches, steps, 2, n_cpu, REF_THREADS)
        out['cpu_baseline'] = {'value': BATCH / sec, 'unit': 'edges/s', 'cores': threads, 'kind': 'port',
                               'sample': f'{n_cpu} steps of the same workload on the host (torch-CPU port of the '
                                         f'reference, self.mlp included, {sec * 1e3:.1f} ms/step)',
                               'at_reference_threads': {'value': BATCH / sec3, 'cores': REF_THREADS,
                                                        'note': 'torch.set_num_threads(3), the reference\'s own setting '
                                                                '(train_link_prediction.py:124)'}}
    del graphs, pair_graphs, m, flush
    torch.cuda.empty_cache()
    return out


# ============================================================================= power-law (sharded) workload
def powerlaw_shape(args):
    import dataclasses
    shape = SHAPES['powerlaw']
    if args.pl_nodes:
        shape = dataclasses.replace(shape, num_src=int(args.pl_nodes))
    return shape


def powerlaw_steps(shape, B, n, seed=1234):
    """Replicated on every rank (same seed): (src, dst, t, neg) per step."""
    rng = np.random.default_rng(seed)
    out = []
    for s, d, t in edge_stream(shape, B, n, seed=seed):
        neg = rng.integers(1, shape.num_src + 1, B).astype(np.int64)
        out.append((s, d, t, neg))
    return out


def cpu_port_powerlaw(shape, B, n_warm, n_timed, threads, scale_down, device='cpu'):
    """The reference's op sequence on a down-scaled replica (its eager decay is N-proportional).  device='cpu': the
    CPU baseline.  A CUDA device: the same stock ATen ops on the GPU (what the unmodified reference does with
    --gpu 0: pageable H2D of the ids per call, scatter_add_ with float atomics, batched GEMM), wall clock."""
    import dataclasses
    from oracle.cpu_port import CpuWalkProjection
    torch.set_num_threads(threads)
    small = dataclasses.replace(shape, num_src=max(shape.num_src // scale_down, 1000))
    ref = CpuWalkProjection(small.node_num, shape.dim, shape.num_layer, shape.time_decay_weight, 0.0,
                            not_scale=False, with_mlp=True, seed=0, device=device)
    steps = powerlaw_steps(small, B, n_warm + n_timed)

    def one(st):
        s, d, t, neg = st
        with torch.no_grad():
            pos = ref.pair_wise(s, d)
            ng = ref.pair_wise(s, neg)
            ref.update(s, d, t)
            return float(pos.sum() - ng.sum())

    for st in steps[:n_warm]:
        one(st)
    if device != 'cpu':
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for st in steps[n_warm:]:
        one(st)                      # ends in float(...): the device is idle again when it returns
    return (time.perf_counter() - t0) / n_timed, small.node_num


def run_powerlaw(args, rank, world, device, K, W, sampler):
    import torch.distributed as dist
    from tpnet_b200.sharded import ShardedRandomProjection
    shape = powerlaw_shape(args)
    # weak scaling (default): every GPU brings its own 100,000 edges per step, so the batch of the whole job is
    # pl_batch x N; strong: the same 100,000-edge batch split over the ranks
    weak = args.scaling == 'weak'
    B = args.pl_batch * (world if weak else 1)
    per_edge_B, per_pair_B = algorithmic_bytes(shape)
    warm_n = args.warm_batches if args.warm_batches is not None else PL_WARM
    n_phase = min(K, 8)
    steps = powerlaw_steps(shape, B, warm_n + 2 * (K + W) + n_phase)
    torch.manual_seed(0)
    m = ShardedRandomProjection(node_num=shape.node_num, edge_num=shape.edge_num, dim_factor=shape.dim_factor,
                                num_layer=shape.num_layer, time_decay_weight=shape.time_decay_weight,
                                device=str(device), use_matrix=False, beginning_time=np.float64(0.0),
                                not_scale=False, enforce_dim=-1, decay_mode='lazy', ext_rows=2 * B + 1024,
                                p0='device', state_device=device)
    m = m.to(device)
    m.init_p0_on_device(seed=0)
    for s, d, t, _ in steps[:warm_n]:
        m.update(s, d, t)
    torch.cuda.synchronize()

    def dev_plan(plan):
        g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)     # noqa: E731
        plan.first_rows, plan.second_rows, plan.send_rows = g(plan.first_rows), g(plan.second_rows), g(plan.send_rows)
        return plan

    def stage(st):
        """Resident inputs of one step.  One GPU: the edge batch itself as device tensors (the
        plain edge-batch path).  Sharded: the routing plans, a pure function of the batch."""
        s, d, t, neg = st
        if world == 1:
            g = lambda a: torch.from_numpy(a).to(device)     # noqa: E731
            return dict(src=g(s), dst=g(d), t=g(t), neg=g(neg), t_last=float(t[-1]))
        up, tmsg = m.plan_update(s, d, t)
        return dict(src=s, dst=d, t=t, up=(dev_plan(up), torch.from_numpy(tmsg).to(device)),
                    pos=dev_plan(m.plan_pairs(s, d)), neg=dev_plan(m.plan_pairs(s, neg)))

    res = [stage(st) for st in steps[warm_n:warm_n + K + W]]

    def resident_pairs(st, which):
        if world == 1:
            m.pair_wise_gram(st['src'], st['dst' if which == 'pos' else 'neg'])
            return B
        m.pair_wise_gram(None, None, plan=st[which])
        return int(st[which].first_rows.shape[0])

    def resident_update(st):
        if world == 1:
            m.update(st['src'], st['dst'], st['t'], next_time=st['t_last'])
            return 2 * B
        m.update(st['src'], st['dst'], st['t'], plan=st['up'])
        return int(st['up'][0].first_rows.shape[0])

    def resident(st):
        resident_pairs(st, 'pos')
        resident_pairs(st, 'neg')
        resident_update(st)

    # One GPU: every step is captured once into a CUDA graph and replayed once, in order (the host-side
    # launch latency of the ~20 small kernels of a step would otherwise show up as idle gaps).
    # Sharded: the all_to_all exchange stays an eager NCCL call.
    use_graphs = world == 1 and not args.no_graphs
    side = torch.cuda.Stream(device)
    pool = torch.cuda.graph_pool_handle() if use_graphs else None

    def capture(fn, *a):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool, stream=side):
            fn(*a)
        return g

    if use_graphs:
        with torch.cuda.stream(side):
            graphs = [capture(resident, st) for st in res]
        torch.cuda.synchronize()
        run_step = lambda i: graphs[i].replay()              # noqa: E731
    else:
        run_step = lambda i: resident(res[i])                # noqa: E731
    for i in range(W):
        run_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    rows0 = m.exchanged_rows
    wall0 = time.perf_counter()
    for k in range(K):
        ev[k][0].record()
        run_step(W + k)
        ev[k][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall1 = time.perf_counter()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    recv_rows_per_step = (m.exchanged_rows - rows0) / K
    m.check_errors()
    del res
    if use_graphs:
        del graphs

    # per-phase device time (events around each call) for the roofline of the dominant kernel
    t_pair, t_upd, n_pairs_local, n_msgs_local = [], [], 0, 0
    for st in steps[warm_n + K + W:warm_n + K + W + n_phase]:
        r = stage(st)
        if world > 1:
            dist.barrier()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        if use_graphs:
            n_pairs_local, n_msgs_local = B, 2 * B
            with torch.cuda.stream(side):
                gp, gu = capture(resident_pairs, r, 'pos'), capture(resident_update, r)
            torch.cuda.synchronize()
            a.record()
            gp.replay()
            b.record()
            gu.replay()
            c.record()
        else:
            a.record()
            n_pairs_local = resident_pairs(r, 'pos')
            b.record()
            n_msgs_local = resident_update(r)
            c.record()
        torch.cuda.synchronize()
        t_pair.append(a.elapsed_time(b))
        t_upd.append(b.elapsed_time(c))
    pair_ms, upd_ms = float(np.mean(t_pair)), float(np.mean(t_upd))

    # end to end: numpy API, plans computed inside the timed region, head + scalar read-back
    api = steps[warm_n + K + W + n_phase:][:K + W]

    def api_one(st):
        s, d, t, neg = st
        with torch.no_grad():
            _, pos = m.get_pair_wise_feature(s, d)
            _, ng = m.get_pair_wise_feature(s, neg)
            m.update(s, d, t)
            r = pos.sum() - ng.sum()
        return float(r.item())

    for st in api[:W]:
        api_one(st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = time.perf_counter()
    n_api = len(api) - W
    for st in api[W:]:
        api_one(st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_ms = (time.perf_counter() - e0) * 1e3
    wall_end = time.perf_counter()

    tt = torch.tensor([dev_ms, e2e_ms, pair_ms, upd_ms, recv_rows_per_step], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, pair_ms, upd_ms, recv_rows_per_step = [float(x) for x in tt.tolist()]
    if rank != 0:
        return None

    peak, peak_src = peak_gbs()
    pair_gbs = n_pairs_local * per_pair_B / (pair_ms * 1e-3) / 1e9
    upd_gbs = n_msgs_local * (per_edge_B / 2) / (upd_ms * 1e-3) / 1e9
    dominant_is_update = upd_ms >= 2 * pair_ms            # two pair-wise calls per step vs one update
    step_bytes = B * per_edge_B + 2 * B * per_pair_B
    block_bytes = (shape.num_layer + 1) * m.row_stride * 4
    state_gb = shape.node_num * block_bytes / 1e9
    roof = {'bound': 'hbm', 'peak': peak, 'unit': 'GB/s', 'peak_source': peak_src,
            'phases': {'pairwise': {'what': 'exchange (N>1) + tpn::pairwise_tma_kernel, %d pairs on rank 0' % n_pairs_local,
                                    'ms': pair_ms, 'achieved': pair_gbs, 'frac': pair_gbs / peak},
                       'update': {'what': 'exchange (N>1) + radix sort + snapshot + tpn::walk_small_kernel || '
                                          'tpn::walk_hub2_kernel, %d messages on rank 0'
                                  % n_msgs_local, 'ms': upd_ms, 'achieved': upd_gbs, 'frac': upd_gbs / peak}}}
    if dominant_is_update:
        roof.update(kernel='update path (radix sort + snapshot + walk_small || walk_hub2)', achieved=upd_gbs,
                    frac=upd_gbs / peak, traffic=traffic_note('powerlaw_update_dram_bytes_per_call'))
    else:
        roof.update(kernel='tpn::pairwise_tma_kernel', achieved=pair_gbs, frac=pair_gbs / peak,
                    traffic=traffic_note('powerlaw_pairwise_dram_bytes_per_launch'))
    h2d = 3 * B * 8 + 2 * (2 * B * 8)
    return {
        'metric': METRIC, 'value': B * K / (dev_ms * 1e-3), 'unit': 'edges/s', 'n_gpus': world, 'steps': K,
        'warmup': W, 'ms_per_step': dev_ms / K, 'higher_is_better': True, 'scaling': 'weak' if weak else 'strong',
        'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': (f'power-law temporal graph, {shape.num_nodes} nodes / {shape.num_edges} edges '
                                f'(BASELINE configs[3]), d={shape.dim}, L={shape.num_layer}, batch {B}, 2 pair-encodes '
                                f'per edge (decoder shape), lazy decay'),
                   'batch': B, 'pairs_per_step': 2 * B, 'dim': shape.dim, 'num_layer': shape.num_layer,
                   'node_num': shape.node_num, 'state_GB': state_gb,
                   'parallelism': 'single GPU' if world == 1 else
                   f'state sharded by node id over {world} GPUs, one all_to_all per call (NCCL)',
                   'l2': f'no flush needed: {state_gb:.1f} GB state >> 126 MB L2',
                   'timing': ('CUDA events per step, one CUDA graph per step' if use_graphs else
                              'CUDA events per step, max over ranks; routing plans are part of the resident inputs'),
                   'algorithmic_bytes_per_step': step_bytes, 'wall_ms_per_step': (wall1 - wall0) * 1e3 / K},
        'pairs_per_s': 2 * B * K / (dev_ms * 1e-3),
        'algorithmic_GBps_step': step_bytes * K / (dev_ms * 1e-3) / 1e9,
        'roofline': roof,
        'e2e': {'value': B * n_api / (e2e_ms * 1e-3), 'unit': 'edges/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': 4, 'ms_per_step': e2e_ms / n_api,
                'path': 'ShardedRandomProjection.get_pair_wise_feature/update with numpy ids (routing plan computed '
                        'on the host inside the timed region), self.mlp included'},
        # per step: 2 pair-wise + update (prep, 3 x (hist + prefix + scatter), payload, giant sort, snapshot,
        # hub2, small, stamps)
        'gpu_launches': K * (2 + 16 + (6 if world > 1 else 0)),
        'exchange': None if world == 1 else {'rows_received_per_rank_per_step': recv_rows_per_step,
                                             'bytes_per_rank_per_step': recv_rows_per_step * block_bytes},
        'clocks': sampler.window(wall0, wall_end) if sampler else None,
    }


# ============================================================================= main
def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    threads = args.cpu_threads or os.cpu_count() or 1

    # ---------------- reference arm: CPU port on the host cores, rank 0 only
    if args.impl == 'reference':
        if rank != 0:
            return
        if args.workload == 'powerlaw':
            K = min(args.steps or 3, 6)
            W = min(args.warmup if args.warmup is not None else 1, 2)
            shape = powerlaw_shape(args)
            sec, n_small = cpu_port_powerlaw(shape, args.pl_batch, W, K, threads, scale_down=10)
            val = args.pl_batch / sec
            sample = (f'{K} batches of {args.pl_batch} edges + 2 pair-encodes per edge on a {n_small}-node replica '
                      f'(10x fewer nodes: the reference\'s eager decay is N-proportional, so this flatters it)')
            cfg = {'workload': f'power-law temporal graph (BASELINE configs[3]) d={shape.dim} L={shape.num_layer} '
                               f'batch {args.pl_batch}, CPU arm on a 10x down-scaled node set', 'batch': args.pl_batch}
        else:
            shape = SHAPES[args.workload]
            K, W = min(args.steps or 30, 60), max(args.warmup or 3, 1)
            warm_batches, steps = make_steps(shape, K + W, seed=0, warm=60)
            sec = cpu_port_tpnet(shape, warm_batches, steps, W, K, threads)
            val = BATCH / sec
            sample = f'{K} steps of the same workload (torch-CPU port of the reference, self.mlp included)'
            cfg = {'workload': f'{shape.name}-shaped synthetic graph, batch {BATCH}, K={NUM_NEIGHBORS}', 'batch': BATCH}
        print(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'edges/s', 'n_gpus': args.gpus,
                          'steps': K, 'warmup': W, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
                          'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
                          'cpu_baseline': {'value': val, 'unit': 'edges/s', 'cores': threads, 'kind': 'port',
                                           'sample': sample},
                          'e2e': {'value': val, 'unit': 'edges/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    # ---------------- our arm
    if not torch.cuda.is_available():
        raise SystemExit('bench.py (impl=ours) needs a CUDA device: there is no CPU fallback')
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    from tpnet_b200 import _lib
    _lib.load()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3 if sampler else 0)

    if args.workload == 'powerlaw':
        K, W = args.steps or 20, max(args.warmup if args.warmup is not None else 3, 3)
        line = run_powerlaw(args, rank, world, device, K, W, sampler)
        if rank == 0:
            if world == 1:
                n_cpu = args.cpu_sample_steps or 3
                sec, n_small = cpu_port_powerlaw(powerlaw_shape(args), args.pl_batch, 1, n_cpu, threads, scale_down=10)
                sec3, _ = cpu_port_powerlaw(powerlaw_shape(args), args.pl_batch, 1, 1, REF_THREADS, scale_down=10)
                line['cpu_baseline'] = {'value': args.pl_batch / sec, 'unit': 'edges/s', 'cores': threads,
                                        'kind': 'port',
                                        'sample': f'{n_cpu} batches on a {n_small}-node replica (10x fewer nodes than '
                                                  f'the GPU run; torch-CPU port of the reference, self.mlp included, '
                                                  f'{sec:.2f} s/step)',
                                        'at_reference_threads': {'value': args.pl_batch / sec3, 'cores': REF_THREADS,
                                                                 'note': 'torch.set_num_threads(3), the reference\'s own '
                                                                         'setting (train_link_prediction.py:124); 1 batch'}}
                if not args.no_also:
                    line['also'] = {'reddit': run_tpnet_shape(args, SHAPES['reddit'], device, 300, 10, with_cpu=True)}
            else:
                line['cpu_baseline'] = None
    else:
        if world > 1:
            raise SystemExit('the TPNet-batch shapes are single-GPU workloads (state 13-32 MB): use --workload powerlaw')
        K, W = args.steps or 300, max(args.warmup if args.warmup is not None else 10, 3)
        t0 = time.perf_counter()
        body = run_tpnet_shape(args, SHAPES[args.workload], device, K, W, with_cpu=True)
        line = {'metric': METRIC, 'n_gpus': 1, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic'}
        line.update(body)
        line['clocks'] = sampler.window(t0, time.perf_counter()) if sampler else None
    if sampler:
        sampler.stop()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
