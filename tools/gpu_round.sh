#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + one full capture of the top kernel.
# Usage (from the repo root, via gpurun): bash tools/gpu_round.sh <tag> [quick]
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
for wl in c2 c3 c5; do
  E=1048576; [ $wl != c2 ] && E=262144
  timeout 900 python bench.py --workload $wl --edges $E --steps 5 --warmup 3 > $OUT/${TAG}_bench_${wl}.json 2> $OUT/${TAG}_bench_${wl}.err; echo "bench $wl rc=$?"
  cat $OUT/${TAG}_bench_${wl}.json
done
if [ "${2:-}" != quick ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
     python bench.py --edges 131072 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k2_true_cost -s 3 -c 1 -o $OUT/${TAG}_k2 -f \
     python bench.py --edges 131072 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
