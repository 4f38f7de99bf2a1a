"""Device-resident batch pipeline (tpnet_b200/pipeline.py): host logic on CPU tensors."""
import numpy as np
import pytest
import torch

from tpnet_b200.pipeline import EpochBatches, replay_updates


def test_batches_are_views_with_the_host_clock():
    rng = np.random.default_rng(0)
    n = 1003
    src, dst = rng.integers(1, 50, n), rng.integers(1, 50, n)
    t = np.sort(rng.random(n) * 1e6)
    eb = EpochBatches(src, dst, t, 200, 'cpu', extra={'edge_ids': np.arange(1, n + 1), 'neg': rng.integers(1, 50, n)})
    assert len(eb) == 6 and eb.src.dtype == torch.int64 and eb.t.dtype == torch.float64
    seen = 0
    for i, b in enumerate(eb):
        assert b.index == i and b.start == seen and len(b) == (200 if i < 5 else 3)
        assert b.src.data_ptr() == eb.src[b.start:].data_ptr()            # a view, not a copy
        assert np.array_equal(b.src.numpy(), src[b.start:b.stop]) and np.array_equal(b.dst.numpy(), dst[b.start:b.stop])
        assert np.array_equal(b.t.numpy(), t[b.start:b.stop])
        assert b.t_last == t[b.stop - 1] and isinstance(b.t_last, float)  # TPNet.py:76: the LAST element
        assert np.array_equal(b.extra['edge_ids'].numpy(), np.arange(b.start + 1, b.stop + 1))
        seen = b.stop
    assert seen == n
    with pytest.raises(IndexError):
        eb.batch(6)
    with pytest.raises(ValueError):
        EpochBatches(src, dst[:-1], t, 200, 'cpu')
    with pytest.raises(ValueError):
        EpochBatches(src, dst, t, 200, 'cpu', extra={'x': np.zeros(3)})


def test_replay_calls_update_with_device_views_and_next_time():
    calls = []

    class Recorder:
        def update(self, s, d, t, next_time=None):
            calls.append((s, d, t, next_time))
    t = np.arange(10, dtype=np.float64) * 2.5
    eb = EpochBatches(np.arange(10), np.arange(10)[::-1].copy(), t, 4, 'cpu')
    replay_updates(Recorder(), eb)
    assert [len(c[0]) for c in calls] == [4, 4, 2]
    assert [c[3] for c in calls] == [7.5, 17.5, 22.5]
    assert all(isinstance(c[0], torch.Tensor) and c[2].dtype == torch.float64 for c in calls)
