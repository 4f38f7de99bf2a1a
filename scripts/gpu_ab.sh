# A/B of a debug flag on the power-law bench: FLAGS="0 4"
mkdir -p gpurun_out
for f in $FLAGS; do
  TPN_DEBUG_FLAGS=$f timeout 600 python bench.py --no-also --cpu-sample-steps 1 > gpurun_out/bench_f$f.json 2> gpurun_out/bench_f$f.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_f$f.json'))
    p=d['roofline']['phases']
    print('flags $f: ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'M edges/s | pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), '| e2e ms', round(d['e2e']['ms_per_step'],3))
except Exception as e: print('flags $f failed', e)
PY
done
