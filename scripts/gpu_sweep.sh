set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for st in 12 20; do
  TPN_HUB2_STAGES=$st timeout 300 python bench.py --no-also --cpu-sample-steps 1 --steps 10 > gpurun_out/bench_st$st.json 2> gpurun_out/bench_st$st.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_st$st.json'))
    p=d['roofline']['phases']
    print('stages $st: ms/step', round(d['ms_per_step'],4), 'pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), 'e2e ms', round(d['e2e']['ms_per_step'],3))
except Exception as e: print('stages $st failed', e)
PY
done
echo skip e2e profile
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_pl.csv python bench.py --steps 2 --warmup 3 --no-also --cpu-sample-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/launches_pl.csv | grep -E "tpn::|launches" | head -30
