# r02 (session 2): the round's reference run at N=1 — default bench line (with `also`), reference arm, launch list,
# ncu --set full of the step's kernels
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "update_prepare or step_graphs or head or fused_front" ) > gpurun_out/r02m_pytest_sub.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02m_pytest_sub.log | cut -c1-200
( time python bench.py > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err ); echo "bench rc=$?"
( time python bench.py --impl reference > gpurun_out/r02m_bench_reference.json 2> gpurun_out/r02m_bench_reference.err ); echo "reference rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02m_bench.json'))
p = d['roofline']['phases']
print('value %.1f M  step %.3f ms  pair %.3f (%.3f)  update %.3f (frac %.3f)  e2e %.3f ms = %.1f M' % (d['value'] / 1e6, d['ms_per_step'], p['pairwise']['ms'], p['pairwise']['frac'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step'], d['e2e']['value'] / 1e6))
print('cpu_baseline', d.get('cpu_baseline'))
for k, v in d.get('also', {}).items():
    if isinstance(v, dict):
        print(k, {kk: v[kk] for kk in ('value', 'ms_per_step') if kk in v}, 'e2e', (v.get('e2e') or {}).get('value'), 'roof', (v.get('roofline') or {}).get('frac'))
r = json.load(open('gpurun_out/r02m_bench_reference.json'))
print('reference arm', r.get('value'), r.get('cpu_baseline'))
PY
tail -3 gpurun_out/r02m_bench.err
Q="--no-also --cpu-sample-steps 1"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02m_launches.csv python bench.py $Q --steps 2 --warmup 3 --no-graphs > gpurun_out/r02m_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/r02m_launches.csv > gpurun_out/r02m_launch_summary.txt 2>&1
grep -E "tpn::|launches" gpurun_out/r02m_launch_summary.txt | head -30 | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none \
  --kernel-name "regex:walk_hub2_kernel|walk_small_kernel|pairwise_tma_kernel|snapshot_kernel|head_tc_kernel|front_kernel|payload_kernel" --launch-skip 70 --launch-count 11 \
  -o gpurun_out/r02m_full -f python bench.py $Q --no-graphs --no-prepare --steps 2 --warmup 3 > gpurun_out/r02m_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r02m_full.ncu-rep --page raw --csv > gpurun_out/r02m_full.raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r02m_full.raw.csv > gpurun_out/r02m_full.pick.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|dram__bytes|dram_throughput" gpurun_out/r02m_full.pick.txt | cut -c1-150
rm -f gpurun_out/r02m_full.ncu-rep
