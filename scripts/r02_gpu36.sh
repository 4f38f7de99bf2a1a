# r02 (session 3): GPU suite + default bench line + reference arm + ncu of the streamed-giant kernels on the hub-rank probe
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/r03d_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r03d_pytest_gpu.log | cut -c1-300
( time python bench.py > gpurun_out/r03d_bench.json 2> gpurun_out/r03d_bench.err ); echo "bench rc=$?"
( time python bench.py --impl reference > gpurun_out/r03d_bench_reference.json 2> gpurun_out/r03d_bench_reference.err ); echo "reference rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r03d_bench.json'))
p = d['roofline']['phases']
print('value %.1f M  step %.3f ms (frac %.3f)  pair %.3f (%.3f)  update %.3f (frac %.3f)  e2e %.3f ms = %.1f M' % (d['value'] / 1e6, d['ms_per_step'], p['step']['frac'], p['pairwise']['ms'], p['pairwise']['frac'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step'], d['e2e']['value'] / 1e6))
print('roofline top', {k: d['roofline'][k] for k in ('kernel', 'achieved', 'frac', 'traffic')})
print('cpu_baseline', d.get('cpu_baseline'))
for k, v in d.get('also', {}).items():
    if isinstance(v, dict):
        print(k, {kk: v[kk] for kk in ('value', 'ms_per_step') if kk in v}, 'e2e', (v.get('e2e') or {}).get('value'), 'roof', (v.get('roofline') or {}).get('frac'))
r = json.load(open('gpurun_out/r03d_bench_reference.json'))
print('reference arm', r.get('value'), r.get('cpu_baseline'))
PY
tail -3 gpurun_out/r03d_bench.err
PROBE_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name "regex:walk_stream_kernel|walk_hub2_kernel" --launch-skip 2 --launch-count 2 -o gpurun_out/r03d_stream_full -f python scripts/hub_rank_probe.py > gpurun_out/r03d_stream_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r03d_stream_full.ncu-rep --page raw --csv > gpurun_out/r03d_stream_full.raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r03d_stream_full.raw.csv > gpurun_out/r03d_stream_full.pick.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|dram__bytes|dram_throughput|registers|warps_active" gpurun_out/r03d_stream_full.pick.txt | cut -c1-150
rm -f gpurun_out/r03d_stream_full.ncu-rep
