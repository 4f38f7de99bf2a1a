set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_apauc.py ) > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest_gpu.log
tail -25 gpurun_out/r02g_pytest_gpu.log | cut -c1-200
Q="--no-also --cpu-sample-steps 1 --steps 10"
python bench.py $Q > gpurun_out/r02g_ab_stream.json 2> gpurun_out/r02g_ab_stream.err; echo "stream rc=$?"
TPN_DEBUG_FLAGS=16 python bench.py $Q > gpurun_out/r02g_ab_nostream.json 2> gpurun_out/r02g_ab_nostream.err; echo "nostream rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02g_ab_*.json')):
    try:
        d = json.load(open(f))
        p = d['roofline']['phases']
        print(f.split('r02g_ab_')[1], 'step %.3f ms  pair %.3f  update %.3f  e2e %.3f' % (d['ms_per_step'], p['pairwise']['ms'], p['update']['ms'], d['e2e']['ms_per_step']))
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02g_launches.csv python bench.py $Q --steps 2 --warmup 3 --no-graphs > gpurun_out/r02g_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/r02g_launches.csv > gpurun_out/r02g_launch_summary.txt 2>&1
grep -E "tpn::|launches" gpurun_out/r02g_launch_summary.txt | head -30 | cut -c1-220
