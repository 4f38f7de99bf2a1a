# r02 (session 3): N=8 weak scaling, reference order: streamed giants (default rule) vs the gathering hub walker
N=8
set -x
mkdir -p gpurun_out
for F in 0 32; do
  EXTRA=""; if [ $F = 32 ]; then EXTRA="--no-parity"; fi
  ( TPN_DEBUG_FLAGS=$F timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 $EXTRA ) > gpurun_out/r02u_bench_n${N}_f$F.json 2> gpurun_out/r02u_bench_n${N}_f$F.err; echo "bench n$N flags $F rc=$?"
  tail -c 800 gpurun_out/r02u_bench_n${N}_f$F.err
done
python - <<PY
import json
for F in (0, 32):
    try:
        line=[l for l in open('gpurun_out/r02u_bench_n8_f%d.json' % F) if l.startswith('{')][-1]
        d=json.loads(line)
        p=d['roofline']['phases']
        print('flags', F, 'N=8 ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'M edges/s | pair ms', round(p['pairwise']['ms'],4), 'update ms', round(p['update']['ms'],4), '| e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e M edges/s', round(d['e2e']['value']/1e6,1))
        print('   exchange', d['exchange'])
        print('   parity', d.get('parity'))
    except Exception as e: print(F, 'no bench json', e)
PY
