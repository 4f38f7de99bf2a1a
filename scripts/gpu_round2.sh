# one-call GPU round: parity tests, bench (ours + reference arm), ncu launch list, ncu --set full of the walk/pair-wise kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
( time timeout 700 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_default.err
( time timeout 400 python bench.py --impl reference ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_pl.csv python bench.py --steps 2 --warmup 3 --no-also --cpu-sample-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/launches_pl.csv | grep -E "tpn::|launches" | head -30
timeout 700 ncu --set full --import-source on --clock-control none \
  --kernel-name "regex:walk_hub2_kernel|walk_small_kernel|pairwise_tma_kernel|snapshot_kernel" --launch-skip 50 --launch-count 8 \
  -o gpurun_out/r01b_full -f python bench.py --no-also --cpu-sample-steps 1 --steps 2 --warmup 3 > gpurun_out/r01b_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r01b_full.ncu-rep --page raw --csv > gpurun_out/r01b_full.raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r01b_full.raw.csv > gpurun_out/r01b_full.pick.txt 2>&1
ls -la gpurun_out/
