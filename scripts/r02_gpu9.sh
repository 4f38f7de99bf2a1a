set -x
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/peer_probe.py ) > gpurun_out/r02_peer_probe.json 2> gpurun_out/r02_peer_probe.err; echo "probe rc=$?"
cat gpurun_out/r02_peer_probe.json; grep -v "^\[W\|^W1\|\*\*\*" gpurun_out/r02_peer_probe.err | tail -15 | cut -c1-300
