# r02 (session 2): fused prep+sort launch (payload separate again), L2 prefetch warp in the hub walker, update_prepare
# (state-independent half of the update overlapped with the pair-wise calls): parity subset + A/B bench lines + launch list
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu -k "fused_front or update_prepare or hub_walker or powerlaw_replica or step_graphs or chunked or peer or routing or baseline_shape or equal_timestamps" ) > gpurun_out/r02k_pytest_sub.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02k_pytest_sub.log | cut -c1-200
Q="--no-also --cpu-sample-steps 1 --steps 10"
python bench.py $Q > gpurun_out/r02k_ab_default.json 2> gpurun_out/r02k_ab_default.err; echo "default rc=$?"
python bench.py $Q --no-prepare > gpurun_out/r02k_ab_noprepare.json 2> gpurun_out/r02k_ab_noprepare.err; echo "noprepare rc=$?"
TPN_DEBUG_FLAGS=16 python bench.py $Q --no-prepare > gpurun_out/r02k_ab_legacyfront_noprepare.json 2> gpurun_out/r02k_ab_legacyfront.err; echo "legacy rc=$?"
TPN_DEBUG_FLAGS=32 python bench.py $Q --no-prepare > gpurun_out/r02k_ab_noprefetch_noprepare.json 2> gpurun_out/r02k_ab_noprefetch.err; echo "noprefetch rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02k_ab_*.json')):
    try:
        d = json.load(open(f))
        p = d['roofline']['phases']
        print(f.split('r02k_ab_')[1], 'value %.1f M  step %.3f ms  pair %.3f  update %.3f (frac %.3f)  e2e %.3f ms' % (d['value'] / 1e6, d['ms_per_step'], p['pairwise']['ms'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step']))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/r02k_ab_default.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02k_launches.csv python bench.py $Q --steps 2 --warmup 3 --no-graphs --no-prepare > gpurun_out/r02k_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/r02k_launches.csv > gpurun_out/r02k_launch_summary.txt 2>&1
grep -E "tpn::|launches" gpurun_out/r02k_launch_summary.txt | head -30 | cut -c1-200
