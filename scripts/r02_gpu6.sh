set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "head" ) > gpurun_out/r02f_pytest_head.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02f_pytest_head.log | cut -c1-220
( timeout 300 python scripts/head_bench.py ) > gpurun_out/r02f_head_bench.json 2> gpurun_out/r02f_head_bench.err; echo "bench rc=$?"
cat gpurun_out/r02f_head_bench.json
bash scripts/r02_multi.sh 2
