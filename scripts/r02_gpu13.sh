# r02: AP/AUC parity of the unmodified reference scripts on the drop-in, full Reddit and Flights shapes (1 epoch),
# then the full configs[4] micro-benchmark sweep (160 points)
set -x
mkdir -p gpurun_out
( time timeout 2400 python scripts/apauc_parity.py --shape reddit --epochs 1 --out gpurun_out/r02_apauc_reddit.json --timeout 1500 ) > gpurun_out/r02_apauc_reddit.log 2>&1; echo "apauc reddit rc=$?"
tail -12 gpurun_out/r02_apauc_reddit.log
( time timeout 3000 python scripts/apauc_parity.py --shape flights --epochs 1 --negative historical --out gpurun_out/r02_apauc_flights.json --timeout 2400 ) > gpurun_out/r02_apauc_flights.log 2>&1; echo "apauc flights rc=$?"
tail -12 gpurun_out/r02_apauc_flights.log
( time timeout 1500 python scripts/sweep_update.py ) > gpurun_out/r02_sweep_update_full.jsonl 2> gpurun_out/r02_sweep.err; echo "sweep rc=$?"
wc -l gpurun_out/r02_sweep_update_full.jsonl; tail -2 gpurun_out/r02_sweep_update_full.jsonl | cut -c1-300
