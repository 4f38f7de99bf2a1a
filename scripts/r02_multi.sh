# r02 multi-GPU round (N = $1): sharded == single-GPU on real ranks (NVLink pulls, flag barriers, NCCL plane), then
# the power-law bench at N GPUs (peer data plane in CUDA graphs, parity leg) with the reference and the chunked order
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
nvidia-smi topo -m | head -12
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py ) > gpurun_out/r02_dist_check_n$N.log 2>&1; echo "dist check rc=$?"
tail -25 gpurun_out/r02_dist_check_n$N.log | cut -c1-200
for acc in reference chunked; do
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --accumulation $acc ) > gpurun_out/r02_bench_n${N}_$acc.json 2> gpurun_out/r02_bench_n${N}_$acc.err; echo "bench n$N $acc rc=$?"
  tail -c 1500 gpurun_out/r02_bench_n${N}_$acc.err
done
python - <<PY
import json
for acc in ('reference', 'chunked'):
    try:
        line=[l for l in open('gpurun_out/r02_bench_n${N}_%s.json' % acc) if l.startswith('{')][-1]
        d=json.loads(line)
        p=d['roofline']['phases']
        print(acc, 'N=${N} ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'M edges/s | pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), '| e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e M edges/s', round(d['e2e']['value']/1e6,1))
        print('   exchange', d['exchange'])
        print('   parity', d.get('parity'))
    except Exception as e: print(acc, 'no bench json', e)
PY
