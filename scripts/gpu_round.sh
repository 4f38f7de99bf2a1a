# one-call GPU round (final state of the round): parity tests, bench (ours + reference arm), ncu launch list,
# ncu --set full of the walk / pair-wise / snapshot / head kernels.  Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
( time timeout 700 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -c 1000 gpurun_out/bench_default.err
( time timeout 400 python bench.py --impl reference ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_pl.csv python bench.py --steps 2 --warmup 3 --no-also --no-graphs --cpu-sample-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/launches_pl.csv | grep -E "tpn::|launches" | head -30
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_reddit.csv python bench.py --workload reddit --steps 6 --warmup 3 --cpu-sample-steps 1 > gpurun_out/ncu_bench_reddit.log 2>&1; echo "ncu reddit rc=$?"
python profiles/launch_summary.py gpurun_out/launches_reddit.csv | grep -E "tpn::|launches" | head -12
timeout 700 ncu --set full --import-source on --clock-control none \
  --kernel-name "regex:walk_hub2_kernel|walk_small_kernel|pairwise_tma_kernel|snapshot_kernel|head_forward_kernel" --launch-skip 56 --launch-count 9 \
  -o gpurun_out/r01c_full -f python bench.py --no-also --no-graphs --cpu-sample-steps 1 --steps 2 --warmup 3 > gpurun_out/r01c_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r01c_full.ncu-rep --page raw --csv > gpurun_out/r01c_full.raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r01c_full.raw.csv > gpurun_out/r01c_full.pick.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|dram__bytes" gpurun_out/r01c_full.pick.txt | cut -c1-140
ls -la gpurun_out/ | head -30
