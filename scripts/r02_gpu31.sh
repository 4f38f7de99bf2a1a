# r02 (session 3): streamed giants as their own launch (walk_stream_kernel), hub walker back to its tuned code
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu -k "hub_walker or peer_data_plane or power or step_graphs or update_prepare" 2>&1 | tail -2
for F in 0 32; do
  echo "== TPN_DEBUG_FLAGS=$F"
  TPN_DEBUG_FLAGS=$F timeout 200 python scripts/hub_rank_probe.py 2>&1 | tail -2 | tee gpurun_out/r02y_probe_f$F.txt
done
TPN_DEBUG_FLAGS=0 PROBE_REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02y_launches_f0.csv python scripts/hub_rank_probe.py > gpurun_out/r02y_ncu_f0.log 2>&1
python profiles/launch_summary.py gpurun_out/r02y_launches_f0.csv 2>&1 | grep -E "walk_" | cut -c1-150
for i in 1 2; do
timeout 300 python bench.py --no-also --no-cpu --steps 10 > gpurun_out/r02y_bench$i.json 2> gpurun_out/r02y_bench$i.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open('gpurun_out/r02y_bench$i.json'))
p = d['roofline']['phases']
print('N=1 bench: value %.1f M  step %.3f ms  pair %.3f  update %.3f (frac %.3f)  e2e %.3f ms' % (d['value'] / 1e6, d['ms_per_step'], p['pairwise']['ms'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step']))
PY
done
