"""Device-resident batch pipeline.

The reference converts every batch's id / timestamp arrays to tensors and copies them to the device
inside the batch loop (``train_link_prediction.py:262-276``, ``models/TPNet.py:74-77``).  ``EpochBatches``
uploads the arrays of a whole split ONCE and hands out per-batch views that the drop-in module accepts
directly (``update(src, dst, t, next_time=...)``, ``get_pair_wise_feature(a, b)``,
``get_neighbor_pair_wise_feature(...)`` all take int64 / float64 CUDA tensors): no per-batch host->device
copy, no per-batch synchronisation — ``t_last`` (which the host needs for the f64 decay factors,
TPNet.py:84-85) is read from the host copy of the timestamps.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterator, Optional, Union

import numpy as np
import torch


@dataclass
class Batch:
    index: int
    start: int
    stop: int
    src: torch.Tensor            # int64 [B] view on the device
    dst: torch.Tensor            # int64 [B]
    t: torch.Tensor              # float64 [B]
    t_last: float                # node_interact_times[-1] of the batch (TPNet.py:76), from the host copy
    extra: Dict[str, torch.Tensor]

    def __len__(self) -> int:
        return self.stop - self.start


class EpochBatches:
    """Batches of ``batch_size`` consecutive interactions of one split, resident on ``device``.
    ``extra``: further per-interaction arrays (edge ids, labels, pre-drawn negatives, ...) sliced alongside."""

    def __init__(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, node_interact_times: np.ndarray,
                 batch_size: int, device: Union[str, torch.device], extra: Optional[Dict[str, np.ndarray]] = None):
        n = len(src_node_ids)
        if len(dst_node_ids) != n or len(node_interact_times) != n:
            raise ValueError('src, dst and time arrays must have the same length')
        if batch_size < 1:
            raise ValueError('batch_size must be positive')
        self.batch_size = int(batch_size)
        self.device = torch.device(device)
        self._t_host = np.ascontiguousarray(node_interact_times, dtype=np.float64)

        def up(a, dtype):
            x = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype))
            if self.device.type == 'cuda':
                x = x.pin_memory()
            return x.to(self.device, non_blocking=True)

        self.src = up(src_node_ids, np.int64)
        self.dst = up(dst_node_ids, np.int64)
        self.t = up(self._t_host, np.float64)
        self.extra = {}
        for k, v in (extra or {}).items():
            if len(v) != n:
                raise ValueError(f'extra array {k!r} has {len(v)} entries for {n} interactions')
            self.extra[k] = up(v, np.asarray(v).dtype)
        self.num_interactions = n

    def __len__(self) -> int:
        return (self.num_interactions + self.batch_size - 1) // self.batch_size

    def batch(self, i: int) -> Batch:
        if not 0 <= i < len(self):
            raise IndexError(i)
        lo, hi = i * self.batch_size, min((i + 1) * self.batch_size, self.num_interactions)
        return Batch(index=i, start=lo, stop=hi, src=self.src[lo:hi], dst=self.dst[lo:hi], t=self.t[lo:hi],
                     t_last=float(self._t_host[hi - 1]), extra={k: v[lo:hi] for k, v in self.extra.items()})

    def __iter__(self) -> Iterator[Batch]:
        for i in range(len(self)):
            yield self.batch(i)


def replay_updates(module, batches: EpochBatches) -> None:
    """``module.update`` over a resident split (e.g. rebuilding the walk state after a reload)."""
    for b in batches:
        module.update(b.src, b.dst, b.t, next_time=b.t_last)
