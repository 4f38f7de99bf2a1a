#!/usr/bin/env python
"""bench.py — temporal edges/sec through the walk-projection hot path (update +
pair-wise encode) on synthetic graphs of the BASELINE.json shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload reddit|wikipedia|flights] [--no-flush]

One "step" = one batch of B=200 temporal edges through the hot path exactly as one
TPNet training/eval batch drives it: two encoder calls of 4*B*K
pairs (positive and negative destinations, models/TPNet.py:313-316), two decoder
calls of B pairs (models/modules.py:112), then `update` (train_link_prediction.py:372).
That is p = 8K+2 = 162 pair-encodes per edge at K=20.

Prints ONE JSON line (see the driver contract).  `value` = edges/s with all inputs
resident in HBM (each step replayed as a CUDA graph, L2 flushed between steps,
device-event timing); `e2e` = the same steps through the module's public numpy API
(pinned H2D of ids inside the timed region, `self.mlp` included, a scalar result read
back every step).  `--impl reference` times the CPU port of the reference
(oracle/cpu_port.py) on the host cores instead.
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

from tpnet_b200.synth import SHAPES, RecentNeighbors, edge_stream, tpnet_pair_lists  # noqa: E402

BATCH = 200
NUM_NEIGHBORS = 20
WARM_BATCHES = 400          # untimed: populates P_1..P_L and the neighbour table


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=500)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='reddit', choices=['reddit', 'wikipedia', 'flights'])
    ap.add_argument('--decay-mode', default='auto', choices=['auto', 'eager', 'lazy'])
    ap.add_argument('--no-flush', action='store_true', help='keep L2 warm between steps (reported, not the headline)')
    ap.add_argument('--cpu-sample-steps', type=int, default=30)
    ap.add_argument('--warm-batches', type=int, default=WARM_BATCHES, help='untimed batches that fill the state')
    return ap.parse_args()


# ----------------------------------------------------------------------------- workload
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
            steps.append(dict(src=s, dst=d, t=t, neg=neg, enc_pos=(pa, pb), enc_neg=(na, nb)))
        nbr.insert(s, d)
    return warm_batches, steps


def pairs_per_step():
    return 8 * BATCH * NUM_NEIGHBORS + 2 * BATCH


def algorithmic_bytes(shape):
    L, d = shape.num_layer, shape.dim
    per_edge = 24 * L * d + 24
    per_pair = 2 * (L + 1) * d * 4 + (2 * L + 2) ** 2 * 4 + 16
    return per_edge, per_pair


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


# ----------------------------------------------------------------------------- CPU port (baseline / reference arm)
def cpu_port_run(shape, warm_batches, steps, n_warm, n_timed, threads):
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
    dt = time.perf_counter() - t0
    return dt / n_timed


# ----------------------------------------------------------------------------- ours
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
                enc_pos=tuple(g(a, torch.int64) for a in st['enc_pos']),
                enc_neg=tuple(g(a, torch.int64) for a in st['enc_neg']))


def resident_step(m, ds):
    """One step with device-resident inputs: kernels only (no head, no H2D).  The feature
    tensors are dropped at once: inside a capture their memory returns to the graph pool."""
    m.pair_wise_gram(*ds['enc_pos'])
    m.pair_wise_gram(*ds['enc_neg'])
    m.pair_wise_gram(ds['src'], ds['dst'])
    m.pair_wise_gram(ds['src'], ds['neg'])
    m.update(ds['src'], ds['dst'], ds['t'], next_time=ds['t_last'])


def api_step(m, st):
    """One step through the public numpy API, head included, scalar result read back."""
    with torch.no_grad():
        m.get_pair_wise_feature(*st['enc_pos'])                # encoder features (feed the Mixer in TPNet)
        m.get_pair_wise_feature(*st['enc_neg'])
        pos = m.get_pair_wise_feature(st['src'], st['dst'])    # decoder features -> the step's scalar result
        neg = m.get_pair_wise_feature(st['src'], st['neg'])
        m.update(st['src'], st['dst'], st['t'])
        res = pos.sum() - neg.sum()
    return float(res.item())                                   # D2H read of the step's result


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    shape = SHAPES[args.workload]
    K, W = args.steps, max(args.warmup, 3)
    per_edge_B, per_pair_B = algorithmic_bytes(shape)
    pps = pairs_per_step()
    workload_name = (f'{shape.name}-shaped synthetic graph ({shape.num_nodes} nodes, {shape.num_edges} edges), '
                     f'd={shape.dim}, L={shape.num_layer}, batch {BATCH}, K={NUM_NEIGHBORS}, '
                     f'{pps // BATCH} pair-encodes per edge, random negatives')
    config = {'workload': workload_name, 'batch': BATCH, 'num_neighbors': NUM_NEIGHBORS, 'pairs_per_step': pps,
              'dim': shape.dim, 'num_layer': shape.num_layer, 'node_num': shape.node_num}

    # ---------------- reference arm: CPU port on the host cores, rank 0 only
    if args.impl == 'reference':
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        warm_batches, steps = make_steps(shape, K + W, seed=rank, warm=args.warm_batches)
        sec = cpu_port_run(shape, warm_batches, steps, W, K, threads)
        val = BATCH / sec
        line = {'impl': 'reference', 'metric': 'temporal edges/sec (update+pairwise encode)', 'value': val,
                'unit': 'edges/s', 'n_gpus': args.gpus, 'steps': K, 'warmup': W, 'ms_per_step': sec * 1e3,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': val, 'unit': 'edges/s', 'cores': threads, 'kind': 'port',
                                 'sample': f'{K} steps of the same workload (torch-CPU port of the reference, '
                                           f'self.mlp included)'},
                'e2e': {'value': val, 'unit': 'edges/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
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

    n_steps = 2 * (K + W)
    warm_batches, steps = make_steps(shape, n_steps, seed=rank, warm=args.warm_batches)   # replicas: own stream per rank
    m = build_module(shape, device, args.decay_mode, warm_batches[0][2][0])
    for s, d, t in warm_batches:
        m.update(s, d, t)
    torch.cuda.synchronize()

    # -- device-resident, graph-replayed steps
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
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3 if sampler else 0)
    if world > 1:
        dist.barrier()
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
    if world > 1:
        dist.barrier()
    wall1 = time.perf_counter()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    m.check_errors()

    # -- dominant kernel (the 4BK-pair encoder launch): live CUDA-event timing, L2 flushed.
    #    Each launch is captured alone in a CUDA graph so that no host launch latency sits
    #    between the two events (the flush before it gives the host time to enqueue).
    n_pair = min(K, 100)
    pair_graphs = []
    with torch.cuda.stream(side):
        for k in range(n_pair):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=side):
                m.pair_wise_gram(*dsteps[W + k]['enc_pos'])
            pair_graphs.append(g)
    torch.cuda.synchronize()
    pev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_pair)]
    for k in range(n_pair):
        flush.zero_()
        pev[k][0].record()
        pair_graphs[k].replay()
        pev[k][1].record()
    torch.cuda.synchronize()
    pair_ms_avg = float(np.mean([a.elapsed_time(b) for a, b in pev]))
    pairs_per_launch = len(dsteps[0]['enc_pos'][0])

    # -- end to end through the public numpy API (pinned H2D + head + scalar D2H per step)
    api = steps[K + W:]
    for st in api[:W]:
        api_step(m, st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = time.perf_counter()
    for st in api[W:W + K]:
        api_step(m, st)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - e0) * 1e3
    h2d = 2 * (2 * 4 * BATCH * NUM_NEIGHBORS * 8) + 2 * (2 * BATCH * 8) + 3 * BATCH * 8
    clocks = sampler.window(wall0, time.perf_counter()) if sampler else None
    if sampler:
        sampler.stop()

    # -- max over ranks
    t = torch.tensor([dev_ms, e2e_ms, wall1 - wall0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_s = [float(x) for x in t.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    edges = BATCH * K * world
    value = edges / (dev_ms * 1e-3)
    step_bytes = BATCH * per_edge_B + pps * per_pair_B
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (burst copy)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    achieved = pairs_per_launch * per_pair_B / (pair_ms_avg * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get('pairwise_kernel_dram_bytes_per_launch')
    kernels_per_step = 4 + 1 + shape.num_layer + (0 if m.lazy else 1)
    cpu = None
    try:
        threads = os.cpu_count() or 1
        n_cpu = max(3, args.cpu_sample_steps)
        sec = cpu_port_run(shape, warm_batches, steps, 2, n_cpu, threads)
        cpu = {'value': BATCH / sec, 'unit': 'edges/s', 'cores': threads, 'kind': 'port',
               'sample': f'{n_cpu} steps of the same workload on the host (torch-CPU port of the reference, '
                         f'self.mlp included, {sec * 1e3:.1f} ms/step)'}
    except Exception as exc:  # pragma: no cover
        cpu = {'value': None, 'unit': 'edges/s', 'cores': 0, 'kind': 'port', 'sample': f'failed: {exc}'}

    line = {
        'metric': 'temporal edges/sec (update+pairwise encode)', 'value': value, 'unit': 'edges/s',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': dev_ms / K, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': dict(config, parallelism=('single GPU' if world == 1 else f'{world} independent replicas'),
                       decay_mode='lazy' if m.lazy else 'eager', l2='flushed between steps (256 MiB memset, untimed)'
                       if not args.no_flush else 'warm (no flush)', timing='CUDA events per step, one CUDA graph per step',
                       algorithmic_bytes_per_step=step_bytes, wall_ms_per_step_incl_flush=wall_s * 1e3 / K),
        'pairs_per_s': pps * K * world / (dev_ms * 1e-3),
        'algorithmic_GBps_step': step_bytes * K * world / (dev_ms * 1e-3) / 1e9,
        'roofline': {'bound': 'hbm', 'kernel': 'tpn::pairwise_kernel (4BK-pair encoder launch)', 'achieved': achieved,
                     'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                     'peak_source': peak_src, 'launch_us': pair_ms_avg * 1e3, 'pairs_per_launch': pairs_per_launch,
                     'note': 'state (%.1f MB) is smaller than L2: after the first touch rows are served by L2, so '
                             'algorithmic bytes/s can exceed the HBM copy peak' %
                             (shape.node_num * (shape.num_layer + 1) * m.row_stride * 4 / 1e6)},
        'cpu_baseline': cpu,
        'e2e': {'value': BATCH * K * world / (e2e_ms * 1e-3), 'unit': 'edges/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': 4, 'ms_per_step': e2e_ms / K,
                'path': 'RandomProjectionModule.get_pair_wise_feature/update with numpy ids, self.mlp included'},
        'gpu_launches': kernels_per_step * K,
        'clocks': clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
