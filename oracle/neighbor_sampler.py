"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — numpy restatement of the reference's
`recent` historical-neighbour sampler.

Follows ``/root/reference/utils/utils.py:70-224`` (class NeighborSampler, strategy 'recent'):
  * adjacency built from the edge list in file order, BOTH directions per edge
    (``get_neighbor_sampler`` :239-262), each node's list stably sorted by timestamp      [:107-113]
  * query (node, t): ``i = np.searchsorted(times_of_node, t)`` (left: strictly before t)   [:151]
  * the LAST ``num_neighbors`` of the first i entries, written to the BACK of a zero row    [:211-218]
  * outputs: neighbour ids int64, edge ids int64, times float64, each ``[n, num_neighbors]``.

Parity status: PINNED — ``tests/golden/make_golden.py`` checks it against the imported reference
class on a graph with equal timestamps, repeated edges and queries that sit exactly on interaction
times, and stores the fixture ``tests/golden/sampler_tiny.npz``.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


class RecentNeighborOracle:
    def __init__(self, src: np.ndarray, dst: np.ndarray, edge_ids: np.ndarray, times: np.ndarray, num_nodes: int):
        """``num_nodes`` = max node id + 1 (row 0 is the padding node and stays empty)."""
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        eid = np.asarray(edge_ids, dtype=np.int64)
        t = np.asarray(times, dtype=np.float64)
        # per edge, in file order: (dst, eid, t) appended to src's list, then (src, eid, t) to dst's list [:248-251]
        owner = np.stack([src, dst], axis=1).reshape(-1)
        other = np.stack([dst, src], axis=1).reshape(-1)
        e2 = np.repeat(eid, 2)
        t2 = np.repeat(t, 2)
        order = np.lexsort((np.arange(owner.shape[0]), t2, owner))      # node, then time, ties in insertion order
        self.nbr, self.eid, self.times = other[order], e2[order], t2[order]
        counts = np.bincount(owner, minlength=num_nodes)
        self.offsets = np.zeros(num_nodes + 1, dtype=np.int64)
        np.cumsum(counts, out=self.offsets[1:])
        self.num_nodes = int(num_nodes)

    def get_historical_neighbors(self, node_ids: np.ndarray, node_interact_times: np.ndarray, num_neighbors: int = 20
                                 ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        n, K = len(node_ids), int(num_neighbors)
        out_n = np.zeros((n, K), dtype=np.int64)
        out_e = np.zeros((n, K), dtype=np.int64)
        out_t = np.zeros((n, K), dtype=np.float64)
        for q, (u, tq) in enumerate(zip(node_ids, node_interact_times)):
            lo, hi = self.offsets[u], self.offsets[u + 1]
            i = lo + np.searchsorted(self.times[lo:hi], tq)              # strictly before tq
            take = min(i - lo, K)
            if take > 0:
                out_n[q, K - take:] = self.nbr[i - take:i]
                out_e[q, K - take:] = self.eid[i - take:i]
                out_t[q, K - take:] = self.times[i - take:i]
        return out_n, out_e, out_t
