set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "neighbor" ) > gpurun_out/pytest_nbr.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_nbr.log
( timeout 600 python bench.py --workload reddit --cpu-sample-steps 5 ) > gpurun_out/bench_reddit.json 2> gpurun_out/bench_reddit.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench_reddit.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_reddit.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['launch_us'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
