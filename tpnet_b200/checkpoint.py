"""Packed checkpoints of the walk-projection state.

The reference checkpoints ``model.state_dict()`` (``utils/EarlyStopping.py:64-87``).  Its
``RandomProjectionModule`` is registered under three parents (the backbone, the link predictor and
the model itself), so the L+1 ``[N, d]`` matrices appear three times in that dict — 3 x 33.6 GB for
the 10M-node graph — and a node-sharded state has no single ``state_dict`` at all.  This format
stores the state ONCE, one file per rank, and can be read back into any number of ranks:

    <dir>/<tag>.rank<r>-of-<G>.pt     torch.save of the header
        {'format', 'world', 'rank', 'global_node_num', 'dim', 'num_layer', 'now_time', 'begging_time', 'rows',
         'mlp': state_dict of the trainable head}
    <dir>/<tag>.rank<r>-of-<G>.state  raw little-endian float32 [rows, L+1, dim], C order
                                      (row j of rank r = node j*G + r; pending lazy decay applied)

The state file is written and read in CHUNKS of rows through a bounded host buffer (``chunk_bytes``, default
64 MiB): a 34.6 GB GPU-resident state is saved and restored without ever holding a whole shard in host memory.
(v1 files, which kept the state inside the .pt header, still load.)

``state_dict()`` / ``load_state_dict()`` of the module keep the reference's keys and remain the way
the unmodified reference scripts checkpoint small graphs; this is the path for states that do not
fit one host copy three times, or are sharded.
"""
from __future__ import annotations

import glob
import os
import re
from typing import Dict, List, Tuple

import numpy as np
import torch

FORMAT_V1 = 'tpnet_b200.shard.v1'
FORMAT = 'tpnet_b200.shard.v2'
DEFAULT_CHUNK_BYTES = 64 << 20


def _shard_info(module) -> Tuple[int, int, int, int]:
    """(world, rank, rows owned by this rank, global node count) of a plain or sharded module."""
    world = int(getattr(module, 'world', 1))
    rank = int(getattr(module, 'rank', 0))
    rows = int(getattr(module, 'n_local', module.node_num))
    global_n = int(getattr(module, 'global_node_num', module.node_num))
    return world, rank, rows, global_n


def shard_path(directory: str, tag: str, rank: int, world: int) -> str:
    return os.path.join(directory, f'{tag}.rank{rank}-of-{world}.pt')


def pack_shard(state_rows: torch.Tensor, world: int, rank: int, global_node_num: int, now_time: float,
               begging_time: float, head: Dict[str, torch.Tensor]) -> dict:
    """Payload of one rank.  ``state_rows``: float32 [rows, L+1, dim] of the nodes rank, rank+world, ..."""
    rows_expected = (global_node_num - rank + world - 1) // world if global_node_num > rank else 0
    if state_rows.dim() != 3 or state_rows.shape[0] != rows_expected or state_rows.dtype != torch.float32:
        raise ValueError(f'rank {rank} of {world} owns {rows_expected} of {global_node_num} nodes; got state '
                         f'{tuple(state_rows.shape)} {state_rows.dtype}')
    return {'format': FORMAT_V1, 'world': int(world), 'rank': int(rank), 'global_node_num': int(global_node_num),
            'dim': int(state_rows.shape[2]), 'num_layer': int(state_rows.shape[1]) - 1, 'now_time': float(now_time),
            'begging_time': float(begging_time), 'state': state_rows.detach().to('cpu').contiguous(),
            'mlp': {k: v.detach().to('cpu') for k, v in head.items()}}


def scatter_rows(dst_rows: torch.Tensor, world: int, rank: int, saved_world: int, saved_rank: int, src: torch.Tensor,
                 first_row: int = 0) -> int:
    """``src``: rows [first_row, first_row + len) of the shard saved by ``saved_rank`` of ``saved_world``
    (float32 [n, L+1, dim]).  Copies those that rank ``rank`` of ``world`` owns into ``dst_rows``
    ([rows, L+1, >= dim], local row u // world of node u).  Returns the number of rows copied."""
    d = int(src.shape[2])
    ids = saved_rank + saved_world * torch.arange(first_row, first_row + src.shape[0], dtype=torch.int64)
    mine = (ids % world) == rank
    if bool(mine.any()):
        local = torch.div(ids[mine], world, rounding_mode='floor')
        dst_rows[local.to(dst_rows.device), :, :d] = src[mine].to(dst_rows.device)
    return int(mine.sum())


def scatter_shard(dst_rows: torch.Tensor, world: int, rank: int, payload: dict) -> int:
    """v1 payload (state inside the header): see ``scatter_rows``."""
    return scatter_rows(dst_rows, world, rank, int(payload['world']), int(payload['rank']), payload['state'])


def state_path(header_path: str) -> str:
    return header_path[:-3] + '.state'


def save_checkpoint(module, directory: str, tag: str = 'walk_state', chunk_bytes: int = DEFAULT_CHUNK_BYTES) -> str:
    """Writes this rank's header + state file (every rank of a sharded state calls it).  The state streams to disk
    in chunks of rows through two pinned host buffers (device -> host copy of chunk k+1 overlaps the write of chunk
    k).  Returns the header path."""
    os.makedirs(directory, exist_ok=True)
    module.materialize()                                    # lazy decay: bring every row current first
    world, rank, rows, global_n = _shard_info(module)
    L1, d = module.num_layer + 1, module.dim
    path = shard_path(directory, tag, rank, world)
    spath = state_path(path)
    row_bytes = L1 * d * 4
    per = max(1, int(chunk_bytes) // row_bytes)
    state = module._state
    on_gpu = state.is_cuda
    bufs = [torch.empty(min(per, max(rows, 1)), L1, d, dtype=torch.float32, pin_memory=on_gpu) for _ in range(2)]
    events = [torch.cuda.Event() if on_gpu else None for _ in range(2)]

    def fetch(k: int):
        lo, hi = k * per, min((k + 1) * per, rows)
        bufs[k & 1][:hi - lo].copy_(state[lo:hi, :, :d], non_blocking=on_gpu)
        if on_gpu:
            events[k & 1].record()
        return lo, hi

    nchunks = (rows + per - 1) // per
    with open(spath + '.tmp', 'wb') as fh:
        pending = fetch(0) if nchunks else None
        for k in range(nchunks):
            lo, hi = pending
            if on_gpu:
                events[k & 1].synchronize()
            if k + 1 < nchunks:
                pending = fetch(k + 1)
            fh.write(memoryview(bufs[k & 1][:hi - lo].numpy()).cast('B'))      # leading rows of a contiguous buffer
    header = {'format': FORMAT, 'world': int(world), 'rank': int(rank), 'global_node_num': int(global_n), 'dim': int(d),
              'num_layer': int(module.num_layer), 'now_time': float(module._now_host),
              'begging_time': float(module.begging_time.item()), 'rows': int(rows),
              'mlp': {k: v.detach().to('cpu') for k, v in module.mlp.state_dict().items()}}
    torch.save(header, path + '.tmp')
    os.replace(spath + '.tmp', spath)
    os.replace(path + '.tmp', path)                         # the header appears last: a reader never sees half a shard
    return path


def list_shards(directory: str, tag: str) -> List[str]:
    """The complete, consistent set of shard files of a checkpoint (raises if ranks are missing or mixed)."""
    pat = re.compile(re.escape(tag) + r'\.rank(\d+)-of-(\d+)\.pt$')
    found = {}
    for p in glob.glob(os.path.join(directory, tag + '.rank*-of-*.pt')):
        m = pat.search(os.path.basename(p))
        if m:
            found[(int(m.group(1)), int(m.group(2)))] = p
    worlds = {w for _, w in found}
    if not found:
        raise FileNotFoundError(f'no checkpoint {tag!r} in {directory}')
    if len(worlds) != 1:
        raise ValueError(f'checkpoint {tag!r} in {directory} mixes world sizes {sorted(worlds)}')
    world = worlds.pop()
    missing = [r for r in range(world) if (r, world) not in found]
    if missing:
        raise FileNotFoundError(f'checkpoint {tag!r}: shards of ranks {missing} of {world} are missing')
    return [found[(r, world)] for r in range(world)]


def load_checkpoint(module, directory: str, tag: str = 'walk_state', load_head: bool = True,
                    chunk_bytes: int = DEFAULT_CHUNK_BYTES) -> None:
    """Reads a checkpoint written with ANY number of ranks into this module's rows (re-sharding by node id:
    a single-GPU state can be loaded into 8 shards and back).  State files are memory-mapped and copied in chunks."""
    world, rank, rows, global_n = _shard_info(module)
    total = 0
    header = None
    with torch.no_grad():
        for path in list_shards(directory, tag):
            payload = torch.load(path, map_location='cpu', mmap=True, weights_only=True)
            if payload.get('format') not in (FORMAT, FORMAT_V1):
                raise ValueError(f'{path}: not a {FORMAT} file')
            if (int(payload['global_node_num']) != global_n or int(payload['dim']) != module.dim
                    or int(payload['num_layer']) != module.num_layer):
                raise ValueError(f'{path}: checkpoint is for {payload["global_node_num"]} nodes, dim {payload["dim"]}, '
                                 f'{payload["num_layer"]} layers; the module has {global_n}, {module.dim}, '
                                 f'{module.num_layer}')
            if payload['format'] == FORMAT_V1:
                total += scatter_shard(module._state[:rows], world, rank, payload)
            else:
                n, L1, d = int(payload['rows']), int(payload['num_layer']) + 1, int(payload['dim'])
                if n:
                    want = n * L1 * d * 4
                    if os.path.getsize(state_path(path)) != want:
                        raise ValueError(f'{state_path(path)}: {os.path.getsize(state_path(path))} bytes, expected {want}')
                    mm = np.memmap(state_path(path), dtype='<f4', mode='r', shape=(n, L1, d))
                    per = max(1, int(chunk_bytes) // (L1 * d * 4))
                    for lo in range(0, n, per):
                        chunk = torch.from_numpy(np.array(mm[lo:lo + per]))       # one bounded host copy
                        total += scatter_rows(module._state[:rows], world, rank, int(payload['world']),
                                              int(payload['rank']), chunk, first_row=lo)
                    del mm
            header = payload
        if total != rows:
            raise ValueError(f'checkpoint covers {total} of the {rows} rows of rank {rank}')
        module.now_time.data.fill_(header['now_time'])
        module.begging_time.data.fill_(header['begging_time'])
        if load_head:
            module.mlp.load_state_dict(header['mlp'])
    module._after_external_write()                          # host clock mirror, lazy bookkeeping restart
