set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu -k "streamed or step_graphs or hub_walker_bit_exact or powerlaw_replica or reload or reset or fixture" ) > gpurun_out/r02h_pytest_sub.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02h_pytest_sub.log | cut -c1-200
cat > /tmp/chain_probe.py <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from tpnet_b200 import RandomProjectionModule, _lib
lib = _lib.load()
dev = 'cuda:0'
N, dim, L = 200001, 210, 3
kw = dict(node_num=N, edge_num=10**9, dim_factor=10, num_layer=L, time_decay_weight=1e-7, device=dev, use_matrix=False,
          beginning_time=np.float64(0.0), not_scale=False, enforce_dim=-1)
rng = np.random.default_rng(0)
for hub_msgs in (71000, 286000):
    B = hub_msgs
    s = rng.integers(1, N, B).astype(np.int64); d = rng.integers(1, N, B).astype(np.int64)
    s[: hub_msgs // 2 + 2000] = 5; d[hub_msgs // 2 + 2000:] = 5
    ds, dd = torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)
    for flag, name in ((16, 'hub walker'), (0, 'streamed')):
        torch.manual_seed(0)
        m = RandomProjectionModule(decay_mode='lazy', **kw).to(dev)
        old = lib.tpn_set_debug_flags(flag)
        t = 0.0
        times = []
        for it in range(5):
            ts = torch.from_numpy(np.sort(t + rng.random(B) * 30.0)).to(dev)
            t = float(ts[-1])
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); m.update(ds, dd, ts, next_time=t); b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        lib.tpn_set_debug_flags(old)
        print(f'hub of ~{hub_msgs} messages, {name}: update {np.median(times[1:]):.3f} ms', flush=True)
        del m
PY
python /tmp/chain_probe.py 2>&1 | tail -6
