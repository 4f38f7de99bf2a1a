# 2-GPU bench only (weak scaling)
set -x
mkdir -p gpurun_out
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus ${NGPU:-2} --steps ${STEPS:-10} --warmup 3 ) > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    line=[l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1]
    d=json.loads(line)
    p=d['roofline']['phases']
    print('N=${NGPU:-2}', d['scaling'], 'batch', d['config']['batch'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'M edges/s | pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), '| e2e ms', round(d['e2e']['ms_per_step'],3), 'exchange', d['exchange'])
except Exception as e: print('no bench json', e)
PY
