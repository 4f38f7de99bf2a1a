# r02 (session 3): route_scatter with a plain-load test in front of the compare-and-swap
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r03g_route_launches.csv python scripts/route_probe.py > gpurun_out/r03g_route_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/r03g_route_launches.csv 2>&1 | grep -E "route_|pull_rows" | cut -c1-170 | head -12
