"""bench.py host-side helpers: the accounting the reported numbers rest on (no GPU)."""
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tpnet_b200.synth import SHAPES, edge_stream, tpnet_neighbor_batch, tpnet_pair_lists, RecentNeighbors  # noqa: E402


def test_algorithmic_bytes_are_the_survey_figures():
    """update 24*L*d + 24 B/edge, pair-wise 8*(L+1)*d + 4*(2L+2)^2 + 16 B/pair."""
    want = {'wikipedia': (120, 2, 5784, 3040), 'reddit': (140, 3, 10104, 4752), 'flights': (150, 3, 10824, 5072),
            'powerlaw': (210, 3, 15144, 6992)}
    for name, (d, L, per_edge, per_pair) in want.items():
        shape = SHAPES[name]
        assert (shape.dim, shape.num_layer) == (d, L)
        assert bench.algorithmic_bytes(shape) == (per_edge, per_pair)
    assert bench.pairs_per_step() == 8 * 200 * 20 + 2 * 200 == 32400           # p = 162 pair-encodes per edge


def test_every_rank_generates_the_same_batches():
    """The sharded run replicates the batch: same seed -> same arrays on every rank; ids 1-based, times sorted."""
    shape = dataclasses.replace(SHAPES['powerlaw'], num_src=50_000)
    a = bench.powerlaw_steps(shape, 3000, 3)
    b = bench.powerlaw_steps(shape, 3000, 3)
    assert len(a) == 3
    for (s1, d1, t1, n1), (s2, d2, t2, n2) in zip(a, b):
        assert np.array_equal(s1, s2) and np.array_equal(d1, d2) and np.array_equal(t1, t2) and np.array_equal(n1, n2)
        assert s1.dtype == np.int64 and t1.dtype == np.float64 and s1.min() >= 1 and s1.max() <= shape.num_src
        assert np.all(np.diff(t1) >= 0)
    assert a[0][2][-1] <= a[1][2][0]                                         # batches follow each other in time


def test_structured_encoder_inputs_expand_to_the_reference_pair_lists():
    """tpnet_neighbor_batch (what the bench feeds tpn_pairwise_neighbors) vs tpnet_pair_lists (TPNet.py:313-316)."""
    shape = SHAPES['reddit']
    nbr = RecentNeighbors(shape.node_num, 20)
    for s, d, t in edge_stream(shape, 200, 3, seed=0):
        nbr.insert(s, d)
    neighbours, s2, d2 = tpnet_neighbor_batch(nbr, s, d)
    a, b = tpnet_pair_lists(nbr, s, d)
    assert neighbours.shape == (400, 20) and len(a) == len(b) == 16000
    assert np.array_equal(a, np.tile(neighbours.reshape(-1), 2))
    assert np.array_equal(b, np.concatenate([np.repeat(s2, 20), np.repeat(d2, 20)]))


def test_clock_sampler_window_parsing():
    cs = bench.ClockSampler.__new__(bench.ClockSampler)
    cs.rows = [(10.0, '1965, 1965, 700.1, Not Active, Not Active, Not Active, Not Active'),
               (10.1, '1920, 1965, 950.3, Not Active, Not Active, Not Active, Active'),
               (50.0, '1200, 1965, 100.0, Active, Not Active, Not Active, Not Active')]
    w = cs.window(9.9, 10.2)
    assert w['sm_mhz'] == 1942.5 and w['sm_max_mhz'] == 1965.0 and w['reasons'] == ['sw_power_cap'] and w['samples'] == 2
    assert cs.window(49.9, 50.1)['reasons'] == ['hw_slowdown']
