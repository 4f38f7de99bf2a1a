# r02 call 4: GPU suite; launch lists (reference / chunked accumulation); A/B of snapshot-with-P0 and giant slice widths
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_apauc.py ) > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest_gpu.log
tail -25 gpurun_out/r02d_pytest_gpu.log | cut -c1-200
Q="--no-also --cpu-sample-steps 1 --steps 10"
for acc in reference chunked; do
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02d_launches_$acc.csv python bench.py $Q --steps 2 --warmup 3 --no-graphs --accumulation $acc > gpurun_out/r02d_ncu_$acc.log 2>&1; echo "ncu $acc rc=$?"
  python profiles/launch_summary.py gpurun_out/r02d_launches_$acc.csv > gpurun_out/r02d_launch_summary_$acc.txt 2>&1
  grep -E "tpn::|launches" gpurun_out/r02d_launch_summary_$acc.txt | head -30
done
python bench.py $Q > gpurun_out/r02d_ab_default.json 2> gpurun_out/r02d_ab_default.err; echo "default rc=$?"
TPN_DEBUG_FLAGS=4 python bench.py $Q > gpurun_out/r02d_ab_snapp0.json 2> gpurun_out/r02d_ab_snapp0.err; echo "snapp0 rc=$?"
TPN_DEBUG_FLAGS=4 python bench.py $Q --accumulation chunked > gpurun_out/r02d_ab_snapp0_chunked.json 2> gpurun_out/r02d_ab_snapp0_chunked.err; echo "snapp0 chunked rc=$?"
cp tpnet_b200/_C/libtpnet_b200.so /tmp/lib_default.so
for F in 8 32; do
  TPN_EXTRA_NVCC_FLAGS=-DTPN_HUB2_GIANT_FLOATS=$F python -m tpnet_b200.build --force > /dev/null 2>&1; echo "build $F rc=$?"
  TPN_EXTRA_NVCC_FLAGS=-DTPN_HUB2_GIANT_FLOATS=$F python bench.py $Q > gpurun_out/r02d_ab_giant$F.json 2> gpurun_out/r02d_ab_giant$F.err; echo "giant$F rc=$?"
  TPN_EXTRA_NVCC_FLAGS=-DTPN_HUB2_GIANT_FLOATS=$F TPN_DEBUG_FLAGS=4 python bench.py $Q > gpurun_out/r02d_ab_giant${F}_snapp0.json 2> gpurun_out/r02d_ab_giant${F}_snapp0.err; echo "giant$F snapp0 rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02d_ab_*.json')):
    try:
        d = json.load(open(f))
        p = d['roofline']['phases']
        print(f.split('r02d_ab_')[1], 'step %.3f ms  pair %.3f  update %.3f  e2e %.3f' % (d['ms_per_step'], p['pairwise']['ms'], p['update']['ms'], d['e2e']['ms_per_step']))
    except Exception as e:
        print(f, 'ERR', e)
PY
