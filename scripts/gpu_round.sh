set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_default.json
( time timeout 600 python bench.py --impl reference ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_powerlaw.csv python bench.py --steps 2 --warmup 3 --no-also --cpu-sample-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
