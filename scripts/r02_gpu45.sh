# r02 (session 3): GPU suite + smoke + N=1 bench line at HEAD
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/r03m_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r03m_pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
( time python bench.py --no-also > gpurun_out/r03m_bench.json 2> gpurun_out/r03m_bench.err ); echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r03m_bench.json'))
p = d['roofline']['phases']
print('value %.1f M  step %.3f ms (frac %.3f)  pair %.3f (%.3f)  update %.3f (frac %.3f)  e2e %.3f ms = %.1f M' % (d['value'] / 1e6, d['ms_per_step'], p['step']['frac'], p['pairwise']['ms'], p['pairwise']['frac'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step'], d['e2e']['value'] / 1e6))
print('gpu_launches', d['gpu_launches'], 'clocks', d['clocks'])
PY
