# r02 (session 3): strong scaling once (the same 100,000-edge batch split over 4 GPUs)
N=4
mkdir -p gpurun_out
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-parity ) > gpurun_out/r03o_bench_n${N}_strong.json 2> gpurun_out/r03o_bench_n${N}_strong.err; echo "bench n$N strong rc=$?"
tail -c 300 gpurun_out/r03o_bench_n${N}_strong.err
python - <<PY
import json
line=[l for l in open('gpurun_out/r03o_bench_n4_strong.json') if l.startswith('{')][-1]
d=json.loads(line)
p=d['roofline']['phases']
print('N=4 strong: batch', d['config']['batch'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'M edges/s | pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), '| e2e M edges/s', round(d['e2e']['value']/1e6,1))
PY
