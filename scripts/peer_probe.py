#!/usr/bin/env python
"""2-rank probe (torchrun): how fast does tpn_pull_rows read random / consecutive row blocks out of the peer's HBM
when the peer buffer is mapped (a) with legacy CUDA IPC (cudaIpcOpenMemHandle on a cudaMalloc allocation) and
(b) through torch's symmetric memory (cuMemCreate / cuMemMap, 2 MiB pages)?  Decides the mapping the sharded data
plane uses.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/peer_probe.py
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpnet_b200 import _lib  # noqa: E402
from tpnet_b200.peer import IpcPeerGroup, PeerBuffer  # noqa: E402

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
lib = _lib.load()
L, rs = 3, 216
node_stride = (L + 1) * rs
n_local, ext = 4_000_000, 120_000          # 13.8 GB of local rows per rank
rows = n_local + ext
nbytes = rows * node_stride * 4
out = {}


def run(label, local_ptr, peer_ptrs):
    table = torch.tensor(peer_ptrs, dtype=torch.int64, device=dev)
    mark = torch.zeros(16, dtype=torch.int32, device=dev)
    ctr = torch.zeros(8, dtype=torch.int32, device=dev)
    need = torch.zeros(ext, dtype=torch.int64, device=dev)
    st = _lib.TpnState()
    st.data, st.num_nodes, st.num_layer, st.dim, st.row_stride, st.node_stride = local_ptr, rows, L, 210, rs, node_stride
    sh = _lib.TpnShard()
    sh.world, sh.rank, sh.global_nodes, sh.num_local_rows, sh.ext_rows = world, rank, world * n_local, n_local, ext
    sh.peer_data, sh.mark, sh.counters, sh.need_nodes = table.data_ptr(), mark.data_ptr(), ctr.data_ptr(), need.data_ptr()
    rng = np.random.default_rng(rank)
    peer = (rank + 1) % world
    for kind in ('random', 'consecutive'):
        for n in (10_000, 100_000):
            r = rng.integers(0, n_local, n) if kind == 'random' else (rng.integers(0, n_local - n) + np.arange(n))
            ids = r.astype(np.int64) * world + peer                  # global ids owned by the peer
            need[:n] = torch.from_numpy(ids).to(dev)
            ctr[0], ctr[1] = n, 0
            torch.cuda.synchronize(); dist.barrier()
            stream = torch.cuda.current_stream().cuda_stream
            for _ in range(2):
                assert lib.tpn_pull_rows(ctypes.byref(st), ctypes.byref(sh), stream) == 0
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                lib.tpn_pull_rows(ctypes.byref(st), ctypes.byref(sh), stream)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            out[f'{label} {kind} {n}'] = {'us': ms * 1e3, 'GBps': n * node_stride * 4 / (ms * 1e-3) / 1e9}
            dist.barrier()


# (a) legacy IPC
buf = PeerBuffer(nbytes, dev)
buf.tensor((rows * node_stride,), torch.float32)[:1000].fill_(1.0)
grp = IpcPeerGroup()
ptrs = grp.exchange('probe', buf)
run('ipc', buf.ptr, ptrs)
torch.cuda.synchronize(); dist.barrier()
grp.close(); dist.barrier()
del buf
torch.cuda.empty_cache()

# (b) torch symmetric memory
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(rows * node_stride, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    t[:1000].fill_(1.0)
    run('symm', t.data_ptr(), [int(p) for p in hdl.buffer_ptrs])
    out['symm_info'] = {'world': hdl.world_size, 'rank': hdl.rank, 'signal_pad': bool(len(hdl.signal_pad_ptrs))}
except Exception as e:                                               # noqa: BLE001
    out['symm_error'] = f'{type(e).__name__}: {e}'
if rank == 0:
    print(json.dumps(out, indent=1))
dist.barrier()
dist.destroy_process_group()
