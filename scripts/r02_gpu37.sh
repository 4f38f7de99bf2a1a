# r02 (session 3): feature-stream test + e2e with the two decoder calls on two streams
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "feature_stream or update_prepare or step_graphs" 2>&1 | tail -2
for V in "" "--no-feature-overlap"; do
timeout 300 python bench.py --no-also --no-cpu --steps 20 $V > gpurun_out/r03e_bench.json 2> gpurun_out/r03e_bench.err; echo "bench $V rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r03e_bench.json'))
p = d['roofline']['phases']
print('value %.1f M  step %.3f ms  update %.3f  e2e %.3f ms = %.1f M edges/s' % (d['value'] / 1e6, d['ms_per_step'], p['update']['ms'], d['e2e']['ms_per_step'], d['e2e']['value'] / 1e6))
PY
tail -2 gpurun_out/r03e_bench.err
done
