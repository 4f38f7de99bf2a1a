# r02 (session 2): fused sort front end (one cooperative launch) — parity subset, A/B against the separate launches,
# uniform-batch update under ncu --set full (DRAM bytes vs algorithmic bytes), launch list of the bench step
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu -k "fused_front or hub_walker_bit_exact or powerlaw_replica or step_graphs or chunked or peer or routing or baseline_shape" ) > gpurun_out/r02j_pytest_sub.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02j_pytest_sub.log | cut -c1-200
Q="--no-also --cpu-sample-steps 1 --steps 10"
python bench.py $Q > gpurun_out/r02j_ab_fused.json 2> gpurun_out/r02j_ab_fused.err; echo "fused rc=$?"
TPN_DEBUG_FLAGS=16 python bench.py $Q > gpurun_out/r02j_ab_legacy.json 2> gpurun_out/r02j_ab_legacy.err; echo "legacy rc=$?"
python bench.py $Q --e2e-sync-read > gpurun_out/r02j_ab_fused_syncread.json 2> gpurun_out/r02j_ab_fused_syncread.err; echo "syncread rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02j_ab_*.json')):
    try:
        d = json.load(open(f))
        p = d['roofline']['phases']
        print(f.split('r02j_ab_')[1], 'value %.1f M  step %.3f ms  pair %.3f  update %.3f (frac %.3f)  e2e %.3f ms' % (d['value'] / 1e6, d['ms_per_step'], p['pairwise']['ms'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step']))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/r02j_ab_fused.err
python scripts/uniform_update.py > gpurun_out/r02j_uniform.json 2> gpurun_out/r02j_uniform.err; cat gpurun_out/r02j_uniform.json
TPN_DEBUG_FLAGS=16 python scripts/uniform_update.py > gpurun_out/r02j_uniform_legacy.json 2>> gpurun_out/r02j_uniform.err; cat gpurun_out/r02j_uniform_legacy.json
timeout 600 ncu --set full --import-source on --clock-control none \
  --kernel-name "regex:walk_hub2_kernel|walk_small_kernel|snapshot_kernel|front_kernel|stamp_targets_kernel" --launch-skip 15 --launch-count 5 \
  -o gpurun_out/r02j_uniform_full -f python scripts/uniform_update.py --calls 5 > gpurun_out/r02j_uniform_ncu.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r02j_uniform_full.ncu-rep --page raw --csv > gpurun_out/r02j_uniform_full.raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r02j_uniform_full.raw.csv > gpurun_out/r02j_uniform_full.pick.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|dram__bytes|dram_throughput" gpurun_out/r02j_uniform_full.pick.txt | cut -c1-150
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02j_launches.csv python bench.py $Q --steps 2 --warmup 3 --no-graphs > gpurun_out/r02j_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/r02j_launches.csv > gpurun_out/r02j_launch_summary.txt 2>&1
grep -E "tpn::|launches" gpurun_out/r02j_launch_summary.txt | head -30 | cut -c1-200
