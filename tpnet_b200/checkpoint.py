"""Packed checkpoints of the walk-projection state.

The reference checkpoints ``model.state_dict()`` (``utils/EarlyStopping.py:64-87``).  Its
``RandomProjectionModule`` is registered under three parents (the backbone, the link predictor and
the model itself), so the L+1 ``[N, d]`` matrices appear three times in that dict — 3 x 33.6 GB for
the 10M-node graph — and a node-sharded state has no single ``state_dict`` at all.  This format
stores the state ONCE, one file per rank, and can be read back into any number of ranks:

    <dir>/<tag>.rank<r>-of-<G>.pt   torch.save of
        {'format', 'world', 'rank', 'global_node_num', 'dim', 'num_layer', 'now_time', 'begging_time',
         'state': float32 [rows, L+1, dim]   (row j of rank r = node j*G + r; pending lazy decay applied),
         'mlp': state_dict of the trainable head}

``state_dict()`` / ``load_state_dict()`` of the module keep the reference's keys and remain the way
the unmodified reference scripts checkpoint small graphs; this is the path for states that do not
fit one host copy three times, or are sharded.
"""
from __future__ import annotations

import glob
import os
import re
from typing import Dict, List, Tuple

import torch

FORMAT = 'tpnet_b200.shard.v1'


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
    return {'format': FORMAT, 'world': int(world), 'rank': int(rank), 'global_node_num': int(global_node_num),
            'dim': int(state_rows.shape[2]), 'num_layer': int(state_rows.shape[1]) - 1, 'now_time': float(now_time),
            'begging_time': float(begging_time), 'state': state_rows.detach().to('cpu').contiguous(),
            'mlp': {k: v.detach().to('cpu') for k, v in head.items()}}


def scatter_shard(dst_rows: torch.Tensor, world: int, rank: int, payload: dict) -> int:
    """Copies from one saved shard the rows that rank ``rank`` of ``world`` owns into ``dst_rows``
    ([rows, L+1, >= dim], local row u // world of node u).  Returns the number of rows copied."""
    saved_world, saved_rank = int(payload['world']), int(payload['rank'])
    src = payload['state']
    d = int(payload['dim'])
    ids = saved_rank + saved_world * torch.arange(src.shape[0], dtype=torch.int64)      # global ids of the saved rows
    mine = (ids % world) == rank
    if bool(mine.any()):
        local = torch.div(ids[mine], world, rounding_mode='floor')
        dst_rows[local.to(dst_rows.device), :, :d] = src[mine].to(dst_rows.device)
    return int(mine.sum())


def save_checkpoint(module, directory: str, tag: str = 'walk_state') -> str:
    """Writes this rank's file (every rank of a sharded state calls it).  Returns the path."""
    os.makedirs(directory, exist_ok=True)
    module.materialize()                                    # lazy decay: bring every row current first
    world, rank, rows, global_n = _shard_info(module)
    state = module._state[:rows, :, :module.dim]
    payload = pack_shard(state, world, rank, global_n, float(module._now_host), float(module.begging_time.item()),
                         module.mlp.state_dict())
    path = shard_path(directory, tag, rank, world)
    tmp = path + '.tmp'
    torch.save(payload, tmp)
    os.replace(tmp, path)                                   # a reader never sees a half-written shard
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


def load_checkpoint(module, directory: str, tag: str = 'walk_state', load_head: bool = True) -> None:
    """Reads a checkpoint written with ANY number of ranks into this module's rows (re-sharding by node id:
    a single-GPU state can be loaded into 8 shards and back)."""
    world, rank, rows, global_n = _shard_info(module)
    total = 0
    header = None
    with torch.no_grad():
        for path in list_shards(directory, tag):
            payload = torch.load(path, map_location='cpu', mmap=True, weights_only=True)
            if payload.get('format') != FORMAT:
                raise ValueError(f'{path}: not a {FORMAT} file')
            if (int(payload['global_node_num']) != global_n or int(payload['dim']) != module.dim
                    or int(payload['num_layer']) != module.num_layer):
                raise ValueError(f'{path}: checkpoint is for {payload["global_node_num"]} nodes, dim {payload["dim"]}, '
                                 f'{payload["num_layer"]} layers; the module has {global_n}, {module.dim}, '
                                 f'{module.num_layer}')
            total += scatter_shard(module._state[:rows], world, rank, payload)
            header = payload
        if total != rows:
            raise ValueError(f'checkpoint covers {total} of the {rows} rows of rank {rank}')
        module.now_time.data.fill_(header['now_time'])
        module.begging_time.data.fill_(header['begging_time'])
        if load_head:
            module.mlp.load_state_dict(header['mlp'])
    module._after_external_write()                          # host clock mirror, lazy bookkeeping restart
