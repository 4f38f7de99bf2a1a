# r02 (session 3): the hub rank of N=8 on one GPU: kernel durations of an update with one 286,000-message target
mkdir -p gpurun_out
for F in 0 32; do
  echo "== TPN_DEBUG_FLAGS=$F"
  TPN_DEBUG_FLAGS=$F timeout 200 python scripts/hub_rank_probe.py 2>&1 | tee gpurun_out/r02v_probe_f$F.txt
  TPN_DEBUG_FLAGS=$F PROBE_REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02v_launches_f$F.csv python scripts/hub_rank_probe.py > gpurun_out/r02v_ncu_f$F.log 2>&1
  python profiles/launch_summary.py gpurun_out/r02v_launches_f$F.csv 2>&1 | grep -E "tpn::" | cut -c1-170
done
