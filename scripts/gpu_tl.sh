TPN_EXTRA_NVCC_FLAGS=-DTPN_HUB2_TIMELINE python -m tpnet_b200.build --force >/dev/null && python scripts/hub_timeline.py 2>&1 | grep -E "consumer:|producer w"
