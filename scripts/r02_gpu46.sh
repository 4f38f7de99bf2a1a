# r02 (session 3): persistent double-buffered pair-wise kernel (pairwise_tma2_kernel) vs the one-tile-per-warp kernel
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair or feature_stream or step_graphs or c_abi" 2>&1 | tail -2
for F in 0 128; do
TPN_DEBUG_FLAGS=$F timeout 200 python bench.py --no-also --no-cpu --steps 20 > gpurun_out/r03n_bench_f$F.json 2> gpurun_out/r03n_bench_f$F.err; echo "bench flags $F rc=$?"
python - <<PY
import json
d = json.load(open('gpurun_out/r03n_bench_f$F.json'))
p = d['roofline']['phases']
print('flags $F: value %.1f M  step %.3f ms  pair %.3f (frac %.3f)  update %.3f  e2e %.3f ms' % (d['value'] / 1e6, d['ms_per_step'], p['pairwise']['ms'], p['pairwise']['frac'], p['update']['ms'], d['e2e']['ms_per_step']))
PY
tail -2 gpurun_out/r03n_bench_f$F.err
done
