# tests + power-law bench + launch list (no reddit, short cpu sample)
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --no-also --cpu-sample-steps 1 ) > gpurun_out/bench_pl.json 2> gpurun_out/bench_pl.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_pl.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_pl.json'))
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    print(json.dumps(d['roofline']['phases']))
except Exception as e: print('no bench json', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pl.csv python bench.py --steps 2 --warmup 3 --no-also --cpu-sample-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/launches_pl.csv | grep -E "tpn::|launches" | head -30
