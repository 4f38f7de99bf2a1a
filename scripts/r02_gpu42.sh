# r02 (session 3): would streaming the N=1 top hub (18 % of the batch) pay with the shipped kernel?  threshold 1/8 vs 1/4
mkdir -p gpurun_out
for F in 128 0; do
TPN_DEBUG_FLAGS=$F timeout 300 python bench.py --no-also --no-cpu --steps 20 > gpurun_out/r03j_bench_f$F.json 2> gpurun_out/r03j_bench_f$F.err; echo "bench flags $F rc=$?"
python - <<PY
import json
d = json.load(open('gpurun_out/r03j_bench_f$F.json'))
p = d['roofline']['phases']
print('flags $F: value %.1f M  step %.3f ms  update %.3f  e2e %.3f ms' % (d['value'] / 1e6, d['ms_per_step'], p['update']['ms'], d['e2e']['ms_per_step']))
PY
done
