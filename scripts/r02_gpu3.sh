# r02 call 3: GPU suite (peer data plane in-process, chunked accumulation), then the default N=1 bench line
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_apauc.py ) > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -30 gpurun_out/r02c_pytest_gpu.log | cut -c1-200
( time timeout 1200 python bench.py ) > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02c_bench.err
head -c 3000 gpurun_out/r02c_bench.json
