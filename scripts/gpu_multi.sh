# 2-GPU round: sharded parity check + power-law bench at N=2 (strong scaling of the same graph)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py ) > gpurun_out/dist_check.log 2>&1; echo "dist check rc=$?"
tail -8 gpurun_out/dist_check.log
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 800 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    line=[l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1]
    d=json.loads(line)
    p=d['roofline']['phases']
    print('N=2 ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'M edges/s | pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), '| e2e ms', round(d['e2e']['ms_per_step'],3), 'exchange', d['exchange'])
except Exception as e: print('no bench json', e)
PY
