"""GPU version of the reference's `recent` historical-neighbour sampler
(``/root/reference/utils/utils.py:70-224`` with ``sample_neighbor_strategy='recent'``).

Same constructor data and the same ``get_historical_neighbors(node_ids, node_interact_times,
num_neighbors)`` call and results as the reference class: for every (node, time) the most recent
``num_neighbors`` interactions strictly before ``time``, written to the back of zero rows — neighbour
ids, edge ids (int64) and times (float64), ``[n, num_neighbors]`` each.  The adjacency lives on the
device as one CSR (built here once, per node stably sorted by timestamp); a query batch is one kernel
(``tpn_sampler_recent``).  With ``as_tensors=True`` the results stay on the device and feed
``RandomProjectionModule.get_neighbor_pair_wise_feature`` without a host round trip.
Only the `recent` strategy exists (what TPNet uses); there is no CPU path.
"""
from __future__ import annotations

from typing import Tuple, Union

import numpy as np
import torch

from . import _lib

ArrayOrTensor = Union[np.ndarray, torch.Tensor]


def build_recent_csr(src: np.ndarray, dst: np.ndarray, eid: np.ndarray, t: np.ndarray, num_nodes: int):
    """(offsets[num_nodes+1], neighbour ids, edge ids, times) of the time-sorted adjacency.
    utils/utils.py:248-251: per edge, (dst, eid, t) joins src's list, then (src, eid, t) joins dst's list;
    :107-113: every list is stably sorted by time."""
    owner = np.stack([src, dst], axis=1).reshape(-1)
    other = np.stack([dst, src], axis=1).reshape(-1)
    e2, t2 = np.repeat(eid, 2), np.repeat(t, 2)
    order = np.lexsort((np.arange(owner.shape[0]), t2, owner))
    offsets = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(owner, minlength=num_nodes), out=offsets[1:])
    return offsets, other[order], e2[order], t2[order]


class RecentNeighborSampler:
    def __init__(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, edge_ids: np.ndarray,
                 node_interact_times: np.ndarray, device: Union[str, torch.device], num_nodes: int = None):
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise RuntimeError('tpnet_b200.RecentNeighborSampler samples on CUDA only (no CPU fallback)')
        src = np.asarray(src_node_ids, dtype=np.int64)
        dst = np.asarray(dst_node_ids, dtype=np.int64)
        eid = np.asarray(edge_ids, dtype=np.int64)
        t = np.asarray(node_interact_times, dtype=np.float64)
        if not (len(src) == len(dst) == len(eid) == len(t)):
            raise ValueError('src, dst, edge id and time arrays must have the same length')
        top = int(max(src.max(initial=0), dst.max(initial=0))) + 1
        self.num_nodes = top if num_nodes is None else int(num_nodes)
        if self.num_nodes < top or (len(src) and min(src.min(), dst.min()) < 0):
            raise IndexError('node id out of range')
        offsets, nbr, eids, times = build_recent_csr(src, dst, eid, t, self.num_nodes)
        self.device = dev
        self._offsets = torch.from_numpy(offsets).to(dev)
        self._nbr = torch.from_numpy(nbr).to(dev)
        self._eid = torch.from_numpy(eids).to(dev)
        self._times = torch.from_numpy(times).to(dev)
        if self._nbr.numel() == 0:                          # keep valid base pointers for an empty graph
            self._nbr = torch.zeros(1, dtype=torch.int64, device=dev)
            self._eid = torch.zeros(1, dtype=torch.int64, device=dev)
            self._times = torch.zeros(1, dtype=torch.float64, device=dev)
        self.sample_neighbor_strategy = 'recent'
        self.seed = None                                    # attributes / methods the reference's callers touch

    def reset_random_state(self) -> None:
        """`recent` sampling is deterministic: nothing to reset (utils/utils.py:316-321 exists for the random strategies)."""

    def _to_device(self, a: ArrayOrTensor, dtype: torch.dtype) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64 if dtype == torch.int64 else np.float64)
                                ).to(self.device)

    def get_historical_neighbors(self, node_ids: ArrayOrTensor, node_interact_times: ArrayOrTensor,
                                 num_neighbors: int = 20, as_tensors: bool = False
                                 ) -> Tuple[ArrayOrTensor, ArrayOrTensor, ArrayOrTensor]:
        """utils/utils.py:160-224 for the `recent` strategy."""
        assert num_neighbors > 0, 'Number of sampled neighbors for each node should be greater than 0!'
        n = int(len(node_ids))
        if len(node_interact_times) != n:
            raise ValueError('node ids and times must have the same length')
        if not isinstance(node_ids, torch.Tensor) and n:
            ids = np.asarray(node_ids)
            if ids.min() < 0 or ids.max() >= self.num_nodes:     # the reference indexes a per-node list (utils.py:177)
                raise IndexError(f'node id out of range for a sampler over {self.num_nodes} nodes')
        qn = self._to_device(node_ids, torch.int64).reshape(-1)
        qt = self._to_device(node_interact_times, torch.float64).reshape(-1)
        out_n = torch.empty(n, num_neighbors, dtype=torch.int64, device=self.device)
        out_e = torch.empty(n, num_neighbors, dtype=torch.int64, device=self.device)
        out_t = torch.empty(n, num_neighbors, dtype=torch.float64, device=self.device)
        if n:
            dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
            stream = torch._C._cuda_getCurrentRawStream(dev_index)
            rc = _lib.load().tpn_sampler_recent(self._offsets.data_ptr(), self._nbr.data_ptr(), self._eid.data_ptr(),
                                                self._times.data_ptr(), self.num_nodes, qn.data_ptr(), qt.data_ptr(), n,
                                                int(num_neighbors), out_n.data_ptr(), out_e.data_ptr(), out_t.data_ptr(),
                                                stream)
            if rc:
                _lib.check(rc, 'tpn_sampler_recent')
        if as_tensors:
            return out_n, out_e, out_t
        return out_n.cpu().numpy(), out_e.cpu().numpy(), out_t.cpu().numpy()


def get_neighbor_sampler(data, sample_neighbor_strategy: str = 'recent', device: Union[str, torch.device] = 'cuda:0',
                         **_unused) -> RecentNeighborSampler:
    """Counterpart of ``utils/utils.py:239-262`` for a reference ``Data`` object (fields src_node_ids,
    dst_node_ids, edge_ids, node_interact_times)."""
    if sample_neighbor_strategy != 'recent':
        raise NotImplementedError("only the 'recent' strategy (the one TPNet uses) has a GPU sampler")
    return RecentNeighborSampler(data.src_node_ids, data.dst_node_ids, data.edge_ids, data.node_interact_times, device)
