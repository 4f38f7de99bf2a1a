# r02 (session 3): N = $1 weak-scaling bench line of the shipped build
N=${1:-4}
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/r03b_bench_n${N}.json 2> gpurun_out/r03b_bench_n${N}.err; echo "bench n$N rc=$?"
tail -c 300 gpurun_out/r03b_bench_n${N}.err
python - <<PY
import json
line=[l for l in open('gpurun_out/r03b_bench_n${N}.json') if l.startswith('{')][-1]
d=json.loads(line)
p=d['roofline']['phases']
print('N=${N} ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'M edges/s | pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), 'step frac', round(p['step']['frac'],3), '| e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e M edges/s', round(d['e2e']['value']/1e6,1))
print('   eager update', d['exchange']['eager_update_ms'])
print('   parity', d.get('parity'))
PY
