# r02 (session 3): consumer timeline of walk_stream_kernel over the whole 286,000-message chain (profiling build)
mkdir -p gpurun_out
TPN_EXTRA_NVCC_FLAGS=-DTPN_HUB2_TIMELINE python -m tpnet_b200.build --force > /dev/null 2>&1; echo "build rc=$?"
PROBE_HUB=286000 timeout 200 python scripts/hub_timeline.py 2>&1 | grep -E "^consumer" | tee gpurun_out/r03k_timeline.txt
