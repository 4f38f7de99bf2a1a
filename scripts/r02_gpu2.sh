# r02 call 2: GPU suite on ABI v10 (chunked accumulation, device-side counts, routed peer data plane in-process)
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_apauc.py ) > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_gpu.log
tail -40 gpurun_out/r02b_pytest_gpu.log
