# r02 (session 3): N=2, every giant streamed (TPN_DEBUG_STREAM_ALL): would a lower threshold than 3/8 pay at N=2?
N=2
mkdir -p gpurun_out
( TPN_DEBUG_FLAGS=64 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-parity ) > gpurun_out/r03c_bench_n${N}.json 2> gpurun_out/r03c_bench_n${N}.err; echo "bench n$N rc=$?"
python - <<PY
import json
line=[l for l in open('gpurun_out/r03c_bench_n${N}.json') if l.startswith('{')][-1]
d=json.loads(line)
p=d['roofline']['phases']
print('N=${N} stream-all ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'M edges/s | pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), '| e2e ms', round(d['e2e']['ms_per_step'],3))
print('   eager update', d['exchange']['eager_update_ms'])
PY
