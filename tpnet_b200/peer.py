"""Peer-visible device memory for the node-sharded state (tpnet_b200/sharded.py).

The sharded data plane reads other ranks' rows straight out of their HBM over NVLink, so every rank's state,
stamps and barrier words must be mapped into every other rank's address space:

  * ``PeerBuffer``      — peer-visible device memory, handed to PyTorch as a tensor.  ``symmetric=True`` (what a
                          one-process-per-GPU job uses): ``torch.distributed._symmetric_memory.empty`` —
                          cuMemCreate / cuMemMap allocations with 2 MiB pages.  Otherwise memory owned by the
                          library (``tpn_peer_alloc``: plain ``cudaMalloc``, exportable with legacy CUDA IPC),
                          viewed through ``__cuda_array_interface__``.
  * ``SymmPeerGroup``   — one process per GPU (torchrun), the default: ``symmetric_memory.rendezvous`` maps every
                          rank's buffer into every other rank and returns plain device pointers.  MEASURED
                          (scripts/peer_probe.py, profiles/r02_peer_probe.json): random 3.4 KB row blocks out of a
                          14 GB peer buffer are pulled at 640 GB/s through this mapping, but only at 171 GB/s
                          through a legacy-IPC mapping of the same buffer (consecutive rows: 647 GB/s on both) —
                          the IPC mapping is TLB-bound for random access.
  * ``IpcPeerGroup``    — the same rendezvous with legacy CUDA IPC handles (``cudaIpcOpenMemHandle``), kept for
                          setups without symmetric-memory support; see the measurement above.
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

    def __init__(self, nbytes: int, device: torch.device, symmetric: bool = False):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('peer-visible memory lives on CUDA devices only')
        self.nbytes = (max(int(nbytes), 16) + 15) // 16 * 16
        self.symmetric = bool(symmetric)
        self._t = None
        if self.symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            self._t = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
            self._t.zero_()
            self.ptr = int(self._t.data_ptr())
            return
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
        if self._t is not None:
            return self._t[:numel * _ITEMSIZE[dtype]].view(dtype).view(*[int(s) for s in shape])
        holder = _ArrayView(self, tuple(int(s) for s in shape), _TYPESTR[dtype])
        return torch.as_tensor(holder, device=self.device)

    def ipc_handle(self) -> bytes:
        buf = (ctypes.c_ubyte * 64)()
        _lib.check(self._lib.tpn_ipc_export(self.ptr, buf), 'tpn_ipc_export')
        return bytes(buf)

    def __del__(self):
        try:
            if getattr(self, '_t', None) is not None:
                self._t = None
                self.ptr = 0
            elif getattr(self, 'ptr', 0):
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


class SymmPeerGroup(PeerGroup):
    """One process per GPU: torch symmetric memory maps every rank's buffer into every other rank (2 MiB pages).
    Buffers must be ``PeerBuffer(..., symmetric=True)`` of the same size on every rank; ``exchange`` is collective."""
    symmetric = True

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self._handles = []

    def exchange(self, name: str, buf: PeerBuffer) -> List[int]:
        import torch.distributed._symmetric_memory as symm_mem
        if buf._t is None:
            raise ValueError('SymmPeerGroup needs PeerBuffer(..., symmetric=True)')
        hdl = symm_mem.rendezvous(buf._t, self.group)
        self._handles.append(hdl)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != buf.ptr:
            raise RuntimeError(f'symmetric-memory rendezvous of "{name}" returned an unexpected pointer table')
        return ptrs

    def barrier(self) -> None:
        self._dist.barrier(group=self.group)

    def close(self) -> None:
        self._handles = []


class IpcPeerGroup(PeerGroup):
    """One process per GPU: pointers of the other ranks' buffers through legacy CUDA IPC (see the module docstring
    for why this is not the default)."""

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
