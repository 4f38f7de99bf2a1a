# r02 (session 3): where a routed pair-wise call of the N=8 job spends its time (all ranks simulated on one GPU)
mkdir -p gpurun_out
timeout 300 python scripts/route_probe.py 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r03f_route_launches.csv python scripts/route_probe.py > gpurun_out/r03f_route_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/r03f_route_launches.csv 2>&1 | grep -E "tpn::" | cut -c1-170 | head -30
