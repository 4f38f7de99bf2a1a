"""Synthetic temporal graphs of the shapes BASELINE.json names (the real datasets are
not available offline).  Host-side numpy only; used by bench.py and the tests.

Shapes follow SURVEY.md §8(d) / the reference's processed-data conventions
(``utils/DataLoader.py:96-135``): node ids are 1-based (0 is the padding node),
bipartite graphs put destinations after sources, timestamps are non-decreasing
float64 seconds.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Iterator, Tuple

import numpy as np


@dataclass(frozen=True)
class GraphShape:
    name: str
    num_src: int           # bipartite: sources 1..num_src ; non-bipartite: all nodes
    num_dst: int           # bipartite: destinations num_src+1..num_src+num_dst ; 0 = non-bipartite
    num_edges: int
    num_layer: int
    time_decay_weight: float
    time_span: float       # seconds covered by the full edge list
    day_quantised: bool    # Flights: integer day stamps -> many equal timestamps per batch
    skew: float            # zipf exponent of endpoint popularity
    dim_factor: int = 10

    @property
    def num_nodes(self) -> int:          # without the padding node
        return self.num_src + self.num_dst

    @property
    def node_num(self) -> int:           # constructor argument of the module (incl. pad id 0)
        return self.num_nodes + 1

    @property
    def edge_num(self) -> int:           # constructor argument (edge feature rows incl. pad row)
        return self.num_edges + 1

    @property
    def dim(self) -> int:                # TPNet.py:30-33
        return min(int(math.log(self.edge_num * 2)) * self.dim_factor, self.node_num)


SHAPES: Dict[str, GraphShape] = {
    # configs[0]: Wikipedia-shaped, 2 projection layers
    'wikipedia': GraphShape('wikipedia', 8227, 1000, 157474, 2, 1e-6, 2.68e6, False, 1.3),
    # configs[1]: Reddit-shaped (the headline 1-GPU workload), default 3 layers
    'reddit': GraphShape('reddit', 10000, 984, 672447, 3, 1e-6, 2.68e6, False, 1.3),
    # configs[2]: Flights-shaped, non-bipartite, day-quantised timestamps
    'flights': GraphShape('flights', 13169, 0, 1927145, 3, 1e-6, 122 * 86400.0, True, 1.3),
    # configs[3]: power-law 10M nodes / 1B edges (state sharded over GPUs when N > 1)
    'powerlaw': GraphShape('powerlaw', 10_000_000, 0, 1_000_000_000, 3, 1e-7, 3.0e7, False, 1.2),
}


def _popularity(rng: np.random.Generator, n: int, count: int, skew: float) -> np.ndarray:
    """`count` draws in [0, n) with a zipf(skew) popularity profile, hubs scattered by a
    fixed multiplicative permutation so that hot rows are not address-adjacent."""
    raw = (rng.zipf(skew, count) - 1) % n
    stride = 2654435761 % n
    while math.gcd(stride, n) != 1:
        stride += 1
    return (raw * stride) % n


def edge_stream(shape: GraphShape, batch: int, num_batches: int, seed: int = 0, start_edge: int = 0
                ) -> Iterator[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """Yields (src, dst, t) batches in chronological order.  The per-edge time step is the
    full graph's (time_span / num_edges), so decay per batch is what the real stream has."""
    rng = np.random.default_rng(seed)
    dt = shape.time_span / shape.num_edges
    e = start_edge
    for _ in range(num_batches):
        src = 1 + _popularity(rng, shape.num_src, batch, shape.skew)
        if shape.num_dst:
            dst = 1 + shape.num_src + _popularity(rng, shape.num_dst, batch, shape.skew)
        else:
            dst = 1 + _popularity(rng, shape.num_src, batch, shape.skew)
        t = (e + np.arange(batch, dtype=np.float64) + rng.random(batch)) * dt
        t.sort()
        if shape.day_quantised:
            t = np.floor(t / 86400.0) * 86400.0
        e += batch
        yield src.astype(np.int64), dst.astype(np.int64), t


class RecentNeighbors:
    """Host-side stand-in for the reference's `recent` NeighborSampler
    (``utils/utils.py:160-224``): the last K neighbours of a node, zero-padded at the
    FRONT (``utils/utils.py:211-219``)."""

    def __init__(self, node_num: int, k: int):
        self.k = k
        self.table = np.zeros((node_num, k), dtype=np.int64)

    def lookup(self, nodes: np.ndarray) -> np.ndarray:
        return self.table[nodes]

    def insert(self, src: np.ndarray, dst: np.ndarray) -> None:
        tab = self.table
        for u, v in zip(src.tolist(), dst.tolist()):
            tab[u, :-1] = tab[u, 1:]
            tab[u, -1] = v
            tab[v, :-1] = tab[v, 1:]
            tab[v, -1] = u


def tpnet_neighbor_batch(nbr: RecentNeighbors, src: np.ndarray, dst: np.ndarray):
    """Inputs of the encoder's structured call for one (src, dst) batch (``models/TPNet.py:206-219``):
    rows = concat(src, dst) (2B of them), their K recent neighbours, and the (src, dst) each row
    is encoded against — ``(neighbours [2B, K], src tiled [2B], dst tiled [2B])``."""
    node_ids = np.concatenate([src, dst])
    return (nbr.lookup(node_ids).astype(np.int64), np.tile(src, 2).astype(np.int64),
            np.tile(dst, 2).astype(np.int64))


def tpnet_pair_lists(nbr: RecentNeighbors, src: np.ndarray, dst: np.ndarray):
    """The (a_ids, b_ids) the reference builds from that batch (``models/TPNet.py:313-316``):
    4*B*K pairs."""
    neighbours, s2, d2 = tpnet_neighbor_batch(nbr, src, dst)
    a = np.tile(neighbours.reshape(-1), 2)
    b = np.concatenate([np.repeat(s2, nbr.k), np.repeat(d2, nbr.k)])
    return a.astype(np.int64), b.astype(np.int64)


def write_processed_dataset(shape: GraphShape, root: str, seed: int = 0, edge_feat_dim: int = 172,
                            node_feat_dim: int = 172, dtype=np.float64) -> str:
    """Writes the shape as a dataset the reference scripts load unchanged
    (``utils/DataLoader.py:96-98``; format of ``preprocess_data/preprocess_data.py:84-117``):

        <root>/processed_data/<name>/ml_<name>.csv       columns ,u,i,ts,label,idx  (ids and idx 1-based)
        <root>/processed_data/<name>/ml_<name>.npy       [E+1, edge_feat_dim], row 0 zero, N(0,1) features
        <root>/processed_data/<name>/ml_<name>_node.npy  [N+1, node_feat_dim] zeros

    Every node id occurs at least once (the loader asserts a gap-free id range, DataLoader.py:136-138).
    Day-quantised shapes (Flights) store the day index: the loader's ``convert_time`` turns it into seconds.
    Returns the dataset directory."""
    name = shape.name
    E = shape.num_edges
    need = max(shape.num_src, shape.num_dst) if shape.num_dst else (shape.num_src + 1) // 2
    if E < need:
        raise ValueError(f'{E} edges cannot touch every one of the {shape.num_nodes} nodes')
    src, dst, t = next(iter(edge_stream(shape, E, 1, seed=seed)))
    rng = np.random.default_rng(seed + 991)
    if shape.num_dst:                                   # bipartite: sources 1..S, destinations S+1..S+D
        src[rng.permutation(E)[:shape.num_src]] = np.arange(1, shape.num_src + 1)
        dst[rng.permutation(E)[:shape.num_dst]] = shape.num_src + np.arange(1, shape.num_dst + 1)
    else:                                               # one id space: cover it with both endpoint columns
        half = (shape.num_src + 1) // 2
        pos = rng.permutation(E)
        src[pos[:half]] = np.arange(1, half + 1)
        dst[pos[:shape.num_src - half]] = np.arange(half + 1, shape.num_src + 1)
    ts = np.floor(t / 86400.0) if shape.day_quantised else t
    out = os.path.join(root, 'processed_data', name)
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f'ml_{name}.csv'), 'w') as fh:
        fh.write(',u,i,ts,label,idx\n')
        for j in range(E):
            fh.write(f'{j},{int(src[j])},{int(dst[j])},{float(ts[j])!r},0.0,{j + 1}\n')
    feats = np.zeros((E + 1, edge_feat_dim), dtype=dtype)
    feats[1:] = rng.standard_normal((E, edge_feat_dim)).astype(dtype)
    np.save(os.path.join(out, f'ml_{name}.npy'), feats)
    np.save(os.path.join(out, f'ml_{name}_node.npy'), np.zeros((shape.num_nodes + 1, node_feat_dim), dtype=dtype))
    return out
