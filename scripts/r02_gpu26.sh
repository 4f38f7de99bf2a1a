# r02 (session 3): full GPU suite + N=1 bench with the relative streaming rule (len > 3/8 of the call's messages)
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02t_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02t_pytest_gpu.log | cut -c1-300
timeout 300 python bench.py --no-also --cpu-sample-steps 1 --steps 10 > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02t_bench.json'))
p = d['roofline']['phases']
print('value %.1f M  step %.3f ms  pair %.3f  update %.3f (frac %.3f)  e2e %.3f ms' % (d['value'] / 1e6, d['ms_per_step'], p['pairwise']['ms'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step']))
PY
