set -x
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/micro/peer_paths.cu -o /tmp/peer_paths && /tmp/peer_paths > gpurun_out/r02_micro_peer_paths.txt 2>&1
cat gpurun_out/r02_micro_peer_paths.txt
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --no-parity ) > gpurun_out/r02_bench_n2_instr.json 2> gpurun_out/r02_bench_n2_instr.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_n2_instr.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n2_instr.json') if l.startswith('{')][-1])
print('ms/step', d['ms_per_step'], d['roofline']['phases']['pairwise']['ms'], d['roofline']['phases']['update']['ms'])
print(d['exchange'])
PY
