# r02 (session 3): streamed giants as ONE launch (ticket roles) + the two decoder calls on two streams
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hub_walker or powerlaw or power_law or step_graphs or update_prepare" ) > gpurun_out/r02o_pytest_sub.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02o_pytest_sub.log | cut -c1-300
Q="--no-also --cpu-sample-steps 1 --steps 10"
timeout 200 python bench.py $Q > gpurun_out/r02o_ab_default.json 2> gpurun_out/r02o_ab_default.err; echo "default rc=$?"
TPN_DEBUG_FLAGS=32 timeout 200 python bench.py $Q > gpurun_out/r02o_ab_nostream.json 2> gpurun_out/r02o_ab_nostream.err; echo "nostream rc=$?"
TPN_DEBUG_FLAGS=64 timeout 200 python bench.py $Q > gpurun_out/r02o_ab_stream8k.json 2> gpurun_out/r02o_ab_stream8k.err; echo "stream8k rc=$?"
timeout 200 python bench.py $Q --no-feature-overlap > gpurun_out/r02o_ab_nooverlap.json 2> gpurun_out/r02o_ab_nooverlap.err; echo "nooverlap rc=$?"
python - <<'PY'
import json
for n in ('default', 'nostream', 'stream8k', 'nooverlap'):
    try:
        d = json.load(open('gpurun_out/r02o_ab_%s.json' % n))
        p = d['roofline']['phases']
        print('%-10s value %.1f M  step %.3f ms  pair %.3f  update %.3f (frac %.3f)  e2e %.3f ms' % (n, d['value'] / 1e6, d['ms_per_step'], p['pairwise']['ms'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step']))
    except Exception as e:
        print(n, 'failed', e)
PY
tail -3 gpurun_out/r02o_ab_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02o_launches.csv python bench.py --no-also --cpu-sample-steps 1 --steps 2 --warmup 3 --no-graphs > gpurun_out/r02o_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/r02o_launches.csv > gpurun_out/r02o_launch_summary.txt 2>&1
grep -E "tpn::|launches" gpurun_out/r02o_launch_summary.txt | head -30 | cut -c1-200
