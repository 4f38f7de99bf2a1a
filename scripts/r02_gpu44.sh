# r02 (session 3): one arrival per stage on the full barrier of walk_stream_kernel (was 32)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu -k "hub_walker or peer_data_plane" 2>&1 | tail -2
TPN_DEBUG_FLAGS=0 timeout 200 python scripts/hub_rank_probe.py 2>&1 | tail -3
TPN_DEBUG_FLAGS=0 PROBE_REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r03l_launches.csv python scripts/hub_rank_probe.py > gpurun_out/r03l_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/r03l_launches.csv 2>&1 | grep -E "walk_" | cut -c1-150
