# r02 call 1: full GPU test suite on the new ABI (incl. the AP/AUC drop-in test), full Wikipedia-shape AP/AUC
# run (2 epochs, both arms + cross-eval), stock-ATen-on-B200 comparison.
set -x
mkdir -p gpurun_out profiles
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -8 gpurun_out/r02_pytest_gpu.log
( time timeout 1500 python scripts/apauc_parity.py --shape wikipedia --epochs 2 --out gpurun_out/r02_apauc_wikipedia.json ) > gpurun_out/r02_apauc_wikipedia.log 2>&1; echo "apauc rc=$?"
tail -30 gpurun_out/r02_apauc_wikipedia.log
( time timeout 600 python scripts/aten_gpu_baseline.py --steps 3 ) > gpurun_out/r02_aten_gpu_baseline.json 2> gpurun_out/r02_aten_gpu_baseline.err; echo "aten rc=$?"
cat gpurun_out/r02_aten_gpu_baseline.json
