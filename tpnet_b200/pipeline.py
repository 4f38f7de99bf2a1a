"""Device-resident batch pipeline: ``EpochBatches`` (a split uploaded once) and
``StepGraphs`` (the per-batch sequence of the hot path captured as one CUDA graph per batch).

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
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Union

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


def tpnet_step(module, batch: Batch, out: Dict[str, torch.Tensor]) -> None:
    """The hot-path calls of one TPNet batch under ``torch.no_grad()`` (evaluate_models_utils.py:64-184): decoder
    features of the positive and the negative pairs, then the update.  ``batch.extra['neg']`` holds the pre-drawn
    negative destinations (the evaluation samplers are seeded, utils/utils.py:339-359: the same negatives every epoch).
    The half of the update that does not write the state (sort by target, work lists, pre-batch snapshot) is started
    first, on the module's side stream, and overlaps the feature kernels; ``update`` then only runs the rest.
    The two feature calls are independent: the second runs on the module's feature stream, so the tensor-core head of
    one overlaps the HBM-bound row gather of the other (same results; forked and joined on the current stream)."""
    module.update_prepare(batch.src, batch.dst, batch.t, next_time=batch.t_last)
    cur = torch.cuda.current_stream(batch.src.device)
    fs = module.feature_stream()
    fs.wait_stream(cur)
    with torch.cuda.stream(fs):
        module.get_pair_wise_feature(batch.src, batch.extra['neg'], out=out['neg'][:len(batch)])
    module.get_pair_wise_feature(batch.src, batch.dst, out=out['pos'][:len(batch)])
    cur.wait_stream(fs)
    module.update(batch.src, batch.dst, batch.t, next_time=batch.t_last)


class StepGraphs:
    """One CUDA graph per batch of a FIXED stream of device-resident batches (a validation / test split, or a
    training epoch with pre-drawn negatives): the per-batch sequence of hot-path calls is captured once, in stream
    order, and replayed every epoch — the host's work per batch is one graph launch instead of ~10-40 kernel
    launches with their argument marshalling (what ``bench.py`` times as `value`).

    * ``step(module, batch, out)`` issues the calls of one batch (default: ``tpnet_step``) and writes its results
      into ``out`` — ONE set of buffers shared by all graphs (``out=`` arguments of the feature calls), read by the
      caller after ``replay(i)`` and before ``replay(i + 1)``.
    * The graphs bake in what the host computes per update (the f64 decay factors of TPNet.py:84-85, the lazy-decay
      epoch), so they are valid for ONE trajectory of the clock: replay them in order, starting from the state the
      module had at capture time (e.g. right after ``reset_random_projections()`` or ``reload_random_projections``
      of the same backup).  ``replay`` enforces the order; ``rewind()`` re-arms the sequence after the caller has
      put the module back into the starting state.
    """

    def __init__(self, module, batches: Sequence, step: Callable = tpnet_step,
                 out: Optional[Dict[str, torch.Tensor]] = None, max_batch: Optional[int] = None):
        """``batches``: ``Batch`` objects (``EpochBatches``) for the default step, or whatever a custom ``step``
        consumes (then pass ``out`` — possibly empty — and ``max_batch``, the largest number of edges per update)."""
        dev = module._require_cuda()
        self.module = module
        self.batches = list(batches)
        if not self.batches:
            raise ValueError('no batches to capture')
        bmax = int(max_batch) if max_batch is not None else max(len(b) for b in self.batches)
        f = module.pair_wise_feature_dim
        self.out = out if out is not None else {
            'pos': torch.empty(bmax, f, dtype=torch.float32, device=dev),
            'neg': torch.empty(bmax, f, dtype=torch.float32, device=dev)}
        h = module._h
        module._c_state()                                    # workspaces / stamps exist before any capture
        start = (h.now, h.epoch, float(h.st.cum_floor) if h.st is not None else 1.0, module.now_time.data.clone())
        self._start = start
        self._after: List[tuple] = []
        self.graphs: List[torch.cuda.CUDAGraph] = []
        pool = torch.cuda.graph_pool_handle()                # graphs replay in capture order: they can share memory
        side = torch.cuda.Stream(dev)
        with torch.no_grad(), torch.cuda.stream(side):
            module._warm_buffers(bmax)                       # nothing is allocated lazily inside a capture
            side.synchronize()
            for b in self.batches:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=side):
                    step(module, b, self.out)
                self.graphs.append(g)
                self._after.append((h.now, h.epoch, float(h.st.cum_floor)))
        torch.cuda.synchronize(dev)
        self._restore(start[:3], start[3])
        self._next = 0

    def _restore(self, host, now_tensor=None) -> None:
        h = self.module._h
        h.now, h.epoch = host[0], host[1]
        if h.st is not None:
            h.st.epoch = host[1]
            h.st.cum_floor = host[2]
        if now_tensor is not None:
            self.module.now_time.data.copy_(now_tensor)

    def __len__(self) -> int:
        return len(self.graphs)

    def replay(self, i: int) -> Dict[str, torch.Tensor]:
        if i != self._next:
            raise RuntimeError(f'StepGraphs replay out of order: batch {i} requested, batch {self._next} is next '
                               f'(the graphs bake in the clock trajectory; call rewind() after restoring the start state)')
        self.graphs[i].replay()
        self._restore(self._after[i])                        # host mirrors follow the device state
        self._next += 1
        return self.out

    def rewind(self) -> None:
        """Call after the module is back in the state it had at capture time (reset / reload of the same backup)."""
        self._restore(self._start[:3], self._start[3])
        self._next = 0
