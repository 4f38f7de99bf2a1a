"""Peer-visible device memory for the node-sharded state (tpnet_b200/sharded.py).

The sharded data plane reads other ranks' rows straight out of their HBM over NVLink, so every rank's state,
stamps and barrier words must be mapped into every other rank's address space:

  * ``PeerBuffer``      — device memory owned by the library (``tpn_peer_alloc``: plain ``cudaMalloc``, so that
                          the whole allocation can be exported with CUDA IPC), handed to PyTorch as a tensor
                          through ``__cuda_array_interface__``; freed when the last tensor view dies.
  * ``IpcPeerGroup``    — one process per GPU (torchrun): the 64-byte IPC handles travel through the process
                          group (``all_gather_object``), each rank opens its peers' handles
                          (``cudaIpcOpenMemHandle`` with lazy peer access) and gets plain device pointers.
  * ``LocalPeerGroup``  — all ranks inside ONE process (several ranks may even share one GPU): the "peer"
                          pointers are simply the other ranks' pointers.  This is how the single-GPU test box
                          runs the complete routed data plane (routing, pulls, barriers, device-side counts);
                          only the NVLink transport itself needs more than one GPU.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _lib

_TYPESTR = {torch.float32: '<f4', torch.int32: '<i4', torch.uint8: '|u1', torch.int64: '<i8', torch.float64: '<f8'}
_ITEMSIZE = {torch.float32: 4, torch.int32: 4, torch.uint8: 1, torch.int64: 8, torch.float64: 8}


class PeerBuffer:
    """Zero-filled device memory from ``tpn_peer_alloc`` on ``device``; ``tensor(shape, dtype)`` views it."""

    def __init__(self, nbytes: int, device: torch.device):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('peer-visible memory lives on CUDA devices only')
        self.nbytes = max(int(nbytes), 16)
        ptr = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.tpn_peer_alloc(ctypes.byref(ptr), self.nbytes), 'tpn_peer_alloc')
        self.ptr = int(ptr.value)

    def tensor(self, shape, dtype: torch.dtype) -> torch.Tensor:
        numel = 1
        for s in shape:
            numel *= int(s)
        if numel * _ITEMSIZE[dtype] > self.nbytes:
            raise ValueError('view larger than the buffer')
        holder = _ArrayView(self, tuple(int(s) for s in shape), _TYPESTR[dtype])
        return torch.as_tensor(holder, device=self.device)

    def ipc_handle(self) -> bytes:
        buf = (ctypes.c_ubyte * 64)()
        _lib.check(self._lib.tpn_ipc_export(self.ptr, buf), 'tpn_ipc_export')
        return bytes(buf)

    def __del__(self):
        try:
            if getattr(self, 'ptr', 0):
                self._lib.tpn_peer_free(self.ptr)
                self.ptr = 0
        except Exception:  # interpreter shutdown
            pass


class _ArrayView:
    """What ``torch.as_tensor`` consumes; keeps the buffer alive as long as the tensor lives."""

    def __init__(self, buf: PeerBuffer, shape, typestr: str):
        self._buf = buf
        self.__cuda_array_interface__ = {'shape': shape, 'typestr': typestr, 'data': (buf.ptr, False), 'version': 2,
                                         'strides': None}


class PeerGroup:
    """Rendezvous of the per-rank buffers of one sharded module."""
    world: int
    rank: int

    def exchange(self, name: str, buf: PeerBuffer) -> List[int]:
        """Device pointers, valid in THIS process, of every rank's buffer ``name`` (index = rank)."""
        raise NotImplementedError

    def barrier(self) -> None:
        """Host-level barrier (set-up only; the data path uses tpn_peer_barrier)."""


class LocalPeerGroup(PeerGroup):
    """All ranks in one process.  ``LocalPeerGroup.create(world)`` returns one view per rank; every rank
    registers its buffers (``publish``) before any of them resolves the table (``exchange``).

    ``host_barriers``: ranks that share ONE GPU cannot use the spinning flag barrier (``tpn_peer_barrier``): a
    kernel of rank A that waits for rank B can sit in front of rank B's work in the same hardware queue.  The
    driver of the ranks orders the phases instead — every rank finishes its reads (``update_begin``, pair-wise
    calls) before any rank writes (``update_end``) — and the module skips the device barrier."""
    host_barriers = True

    def __init__(self, world: int, rank: int, table: Dict[str, List[Optional[int]]]):
        self.world, self.rank, self._table = int(world), int(rank), table

    @staticmethod
    def create(world: int) -> List['LocalPeerGroup']:
        table: Dict[str, List[Optional[int]]] = {}
        return [LocalPeerGroup(world, r, table) for r in range(world)]

    def publish(self, name: str, buf: PeerBuffer) -> None:
        self._table.setdefault(name, [None] * self.world)[self.rank] = buf.ptr

    def exchange(self, name: str, buf: PeerBuffer) -> List[int]:
        self.publish(name, buf)
        ptrs = self._table[name]
        missing = [r for r, p in enumerate(ptrs) if p is None]
        if missing:
            raise RuntimeError(f'ranks {missing} have not published "{name}" yet: construct every rank, then connect()')
        return [int(p) for p in ptrs]


class IpcPeerGroup(PeerGroup):
    """One process per GPU: pointers of the other ranks' buffers through CUDA IPC."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._opened: List[int] = []
        self._lib = _lib.load()

    def exchange(self, name: str, buf: PeerBuffer) -> List[int]:
        handles: List[Optional[bytes]] = [None] * self.world
        self._dist.all_gather_object(handles, (name, buf.ipc_handle()), group=self.group)
        out = []
        with torch.cuda.device(buf.device):
            for r, item in enumerate(handles):
                if r == self.rank:
                    out.append(buf.ptr)
                    continue
                peer_name, handle = item
                if peer_name != name:
                    raise RuntimeError(f'rank {r} exchanged "{peer_name}" while rank {self.rank} exchanged "{name}"')
                ptr = ctypes.c_void_p()
                raw = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
                _lib.check(self._lib.tpn_ipc_open(raw, ctypes.byref(ptr)), f'tpn_ipc_open (buffer "{name}" of rank {r})')
                self._opened.append(int(ptr.value))
                out.append(int(ptr.value))
        return out

    def barrier(self) -> None:
        self._dist.barrier(group=self.group)

    def close(self) -> None:
        for p in self._opened:
            self._lib.tpn_ipc_close(p)
        self._opened = []
