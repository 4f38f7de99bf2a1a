set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -k "hub or bit_exact or oracle or lazy or sharded or full_size or pairwise_large" 2>&1 | tail -4 )
( timeout 600 python bench.py --no-also --cpu-sample-steps 1 ) > gpurun_out/bench_pl.json 2> gpurun_out/bench_pl.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/bench_pl.err
( timeout 600 python bench.py --no-also --cpu-sample-steps 1 --no-graphs ) > gpurun_out/bench_pl_nograph.json 2> gpurun_out/bench_pl_nograph.err; echo "bench nograph rc=$?"
python - <<'PY'
import json
for f in ('bench_pl', 'bench_pl_nograph'):
    try:
        d=json.load(open('gpurun_out/%s.json' % f))
        p=d['roofline']['phases']
        print(f, 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'M edges/s | pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), '| e2e ms', round(d['e2e']['ms_per_step'],3))
    except Exception as e: print(f, 'no bench json', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_pl.csv python bench.py --steps 2 --warmup 3 --no-also --cpu-sample-steps 1 --no-graphs > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/launches_pl.csv | grep -E "tpn::|launches" | head -12
