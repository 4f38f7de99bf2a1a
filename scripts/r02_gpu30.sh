# r02 (session 3): giant chain out of line, one software pipeline per 128-message stage: hub-rank probe + timeline + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hub_walker" 2>&1 | tail -2
for F in 0 32; do
  echo "== TPN_DEBUG_FLAGS=$F"
  TPN_DEBUG_FLAGS=$F timeout 200 python scripts/hub_rank_probe.py 2>&1 | tail -2 | tee gpurun_out/r02x_probe_f$F.txt
  TPN_DEBUG_FLAGS=$F PROBE_REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02x_launches_f$F.csv python scripts/hub_rank_probe.py > gpurun_out/r02x_ncu_f$F.log 2>&1
  python profiles/launch_summary.py gpurun_out/r02x_launches_f$F.csv 2>&1 | grep -E "walk_hub2|walk_small" | cut -c1-150
done
timeout 300 python bench.py --no-also --no-cpu --steps 10 > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02x_bench.json'))
p = d['roofline']['phases']
print('N=1 bench: value %.1f M  step %.3f ms  pair %.3f  update %.3f (frac %.3f)  e2e %.3f ms' % (d['value'] / 1e6, d['ms_per_step'], p['pairwise']['ms'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step']))
PY
TPN_EXTRA_NVCC_FLAGS=-DTPN_HUB2_TIMELINE python -m tpnet_b200.build --force > /dev/null 2>&1; echo "timeline build rc=$?"
for F in 0 32; do
  echo "== timeline, hub 286000, TPN_DEBUG_FLAGS=$F"
  TPN_DEBUG_FLAGS=$F PROBE_HUB=286000 timeout 200 python scripts/hub_timeline.py 2>&1 | grep -E "^consumer" | tee gpurun_out/r02x_timeline_f$F.txt
done
