# r02 (session 3): chain speed of the streamed giants vs the gathering hub walker on crafted batches (scripts/hub_probe.py)
mkdir -p gpurun_out
echo "== streamed (default)"; TPN_DEBUG_FLAGS=0 timeout 300 python scripts/hub_probe.py 2>&1 | tee gpurun_out/r02s_probe_stream.txt
echo "== hub walker (TPN_DEBUG_NO_STREAM)"; TPN_DEBUG_FLAGS=32 timeout 300 python scripts/hub_probe.py 2>&1 | tee gpurun_out/r02s_probe_nostream.txt
