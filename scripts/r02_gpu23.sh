# r02 (session 3): when does streaming the giants pay?  One GPU, growing batch = growing hub (the hub rank of N = 2, 4, 8)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hub_walker" > gpurun_out/r02q_pytest_sub.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02q_pytest_sub.log | cut -c1-300
for B in 100000 200000 400000 800000; do
  for F in 0 32; do
    TPN_DEBUG_FLAGS=$F timeout 300 python bench.py --no-also --no-cpu --steps 6 --pl-batch $B > gpurun_out/r02q_b${B}_f${F}.json 2> gpurun_out/r02q_b${B}_f${F}.err; echo "B=$B F=$F rc=$?"
  done
done
python - <<'PY'
import json
for B in (100000, 200000, 400000, 800000):
    for F in (0, 32):
        try:
            d = json.load(open('gpurun_out/r02q_b%d_f%d.json' % (B, F)))
            p = d['roofline']['phases']
            print('B=%-7d flags=%-2d value %.1f M  step %.3f ms  pair %.3f  update %.3f (frac %.3f)  e2e %.3f ms' % (B, F, d['value'] / 1e6, d['ms_per_step'], p['pairwise']['ms'], p['update']['ms'], p['update']['frac'], d['e2e']['ms_per_step']))
        except Exception as e:
            print(B, F, 'failed', e)
PY
