# r02 (session 2): GPU suite at HEAD, AP/AUC parity on the Reddit shape (full, 1 epoch) and a 400k-edge Flights replica
# (historical negatives), full configs[4] sweep
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest_gpu.log
tail -6 gpurun_out/r02i_pytest_gpu.log | cut -c1-200
( time timeout 1500 python scripts/apauc_parity.py --shape reddit --epochs 1 --out gpurun_out/r02_apauc_reddit.json --timeout 1200 ) > gpurun_out/r02_apauc_reddit.log 2>&1; echo "apauc reddit rc=$?"
tail -8 gpurun_out/r02_apauc_reddit.log | cut -c1-300
( time timeout 1200 python scripts/apauc_parity.py --shape flights --edges 400000 --epochs 1 --negative historical --out gpurun_out/r02_apauc_flights.json --timeout 900 ) > gpurun_out/r02_apauc_flights.log 2>&1; echo "apauc flights rc=$?"
tail -8 gpurun_out/r02_apauc_flights.log | cut -c1-300
( time timeout 900 python scripts/sweep_update.py ) > gpurun_out/r02_sweep_update_full.jsonl 2> gpurun_out/r02_sweep.err; echo "sweep rc=$?"
wc -l gpurun_out/r02_sweep_update_full.jsonl; tail -2 gpurun_out/r02_sweep_update_full.jsonl | cut -c1-300
