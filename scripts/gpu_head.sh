set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_head or neighbor or fixture" 2>&1 | tail -12 )
python scripts/e2e_profile_pl.py 2>&1 | head -8
( timeout 600 python bench.py --no-also --cpu-sample-steps 1 ) > gpurun_out/bench_pl.json 2> gpurun_out/bench_pl.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_pl.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_pl.json'))
    p=d['roofline']['phases']
    print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'M edges/s | pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), '| e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e M edges/s', round(d['e2e']['value']/1e6,1))
except Exception as e: print('no bench json', e)
PY
