# r02 (session 2): pipelined tensor-core head, update_prepare, staging copies on their own stream — full GPU suite,
# head bench, bench A/B lines, host profile of the e2e step
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_apauc.py ) > gpurun_out/r02l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest_gpu.log
tail -6 gpurun_out/r02l_pytest_gpu.log | cut -c1-200
python scripts/head_bench.py > gpurun_out/r02l_head_bench.json 2> gpurun_out/r02l_head_bench.err; cat gpurun_out/r02l_head_bench.json
Q="--no-also --cpu-sample-steps 1 --steps 10"
python bench.py $Q > gpurun_out/r02l_ab_default.json 2> gpurun_out/r02l_ab_default.err; echo "default rc=$?"
python bench.py $Q --no-prepare > gpurun_out/r02l_ab_noprepare.json 2> gpurun_out/r02l_ab_noprepare.err; echo "noprepare rc=$?"
TPN_STAGE_COPY_STREAM=0 python bench.py $Q > gpurun_out/r02l_ab_nocopystream.json 2> gpurun_out/r02l_ab_nocopystream.err; echo "nocopystream rc=$?"
python bench.py $Q --e2e-sync-read > gpurun_out/r02l_ab_syncread.json 2> gpurun_out/r02l_ab_syncread.err; echo "syncread rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02l_ab_*.json')):
    try:
        d = json.load(open(f))
        p = d['roofline']['phases']
        print(f.split('r02l_ab_')[1], 'value %.1f M  step %.3f ms  pair %.3f  update %.3f (frac %.3f)  e2e %.3f ms' % (d['value'] / 1e6, d['ms_per_step'], p['pairwise']['ms'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step']))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/r02l_ab_default.err
PL_NODES=10000000 timeout 600 python scripts/e2e_profile_pl.py > gpurun_out/r02l_e2e_profile.txt 2>&1; head -12 gpurun_out/r02l_e2e_profile.txt
