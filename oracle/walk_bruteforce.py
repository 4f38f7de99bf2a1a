"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — exhaustive enumeration of
temporal walks, the known-answer test the reference ships in
``/root/reference/demo_on_matrix_updating.ipynb`` (cell 2, ``get_matrix_by_brute_force``
with ``matrix_type='sum'``; asserted against the incremental update at cell 8 /
cell 10 with ``rtol=1e-5, atol=1e-5``).

Definition restated (notebook cell 0): a k-step temporal walk from u is a
sequence (u=w_0, t_0=T) -> (w_1, t_1) -> ... -> (w_k, t_k) where every hop
uses an interaction {w_{i-1}, w_i} at time t_i STRICTLY earlier than t_{i-1},
and T = (last timestamp) + 1.  The "sum" walk matrix is
    A^(k)[u, v] = sum over k-step walks u ~> v of  prod_{i=1..k} exp(-lambda (T - t_i)).
A^(0) is the identity.  Pure Python, exponential in k: small graphs only.
"""
from __future__ import annotations

import bisect
import math
from typing import List

import numpy as np


def sum_walk_matrices(src: np.ndarray, dst: np.ndarray, times: np.ndarray, num_layer: int,
                      lam: float, node_num: int) -> List[np.ndarray]:
    """Returns [A^(0), ..., A^(num_layer)] as float64 [node_num, node_num] at time
    T = times[-1] + 1.  ``times`` must be non-decreasing."""
    nbr: List[List[int]] = [[] for _ in range(node_num)]
    when: List[List[float]] = [[] for _ in range(node_num)]
    for u, v, t in zip(src.tolist(), dst.tolist(), times.tolist()):
        nbr[u].append(v); when[u].append(t)
        nbr[v].append(u); when[v].append(t)
    horizon = float(times[-1]) + 1.0
    out = [np.zeros((node_num, node_num), dtype=np.float64) for _ in range(num_layer + 1)]

    def extend(origin: int, node: int, before: float, hops: int, score: float) -> None:
        out[hops][origin, node] += score
        if hops == num_layer:
            return
        upto = bisect.bisect_left(when[node], before)       # interactions strictly earlier
        for k in range(upto):
            t_k = when[node][k]
            extend(origin, nbr[node][k], t_k, hops + 1, score * math.exp(-lam * (horizon - t_k)))

    for u in range(node_num):
        extend(u, u, horizon, 0, 1.0)
    return out


def random_temporal_graph(node_num: int, edge_num: int, rng: np.random.Generator):
    """Random graph in the style of notebook cell 2 ``generate_graph``: distinct
    endpoints, strictly increasing integer timestamps with gaps 1..4."""
    src = rng.integers(0, node_num, size=edge_num)
    off = rng.integers(1, node_num, size=edge_num)
    dst = (src + off) % node_num
    times = np.cumsum(rng.integers(1, 5, size=edge_num)).astype(np.float64)
    return src.astype(np.int64), dst.astype(np.int64), times
