# r02 (session 3): consumer timeline of a streamed giant (profiling build)
mkdir -p gpurun_out
TPN_EXTRA_NVCC_FLAGS=-DTPN_HUB2_TIMELINE python -m tpnet_b200.build --force > /dev/null 2>&1; echo "build rc=$?"
for H in 100000 286000; do
  echo "== hub $H"
  PROBE_HUB=$H timeout 200 python scripts/hub_timeline.py 2>&1 | grep -E "consumer|stage" | tee gpurun_out/r02w_timeline_$H.txt
done
