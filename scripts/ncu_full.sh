# usage: bash scripts/ncu_full.sh <out-name> <kernel-regex> <skip> <count> [bench args...]
set -x
out=$1; regex=$2; skip=$3; count=$4; shift 4
mkdir -p gpurun_out
timeout 1500 ncu --set full --import-source on --clock-control none \
  --kernel-name "regex:$regex" --launch-skip $skip --launch-count $count \
  -o gpurun_out/$out -f python bench.py --no-also --cpu-sample-steps 1 "$@" > gpurun_out/$out.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/$out.ncu-rep --page raw --csv > gpurun_out/$out.raw.csv 2>/dev/null
ls -la gpurun_out/
