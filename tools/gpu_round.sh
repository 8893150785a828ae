#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full captures of the top kernels.
# Usage (from the repo root, via gpurun): bash tools/gpu_round.sh <tag> [quick]
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
for wl in c2 c3 c5; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 > $OUT/${TAG}_bench_${wl}.json 2> $OUT/${TAG}_bench_${wl}.err; echo "bench $wl rc=$?"
  cat $OUT/${TAG}_bench_${wl}.json
done
if [ "${2:-}" != quick ]; then
  # launch list of the default bench command (cold-cache, serialised launches: shares, not absolutes)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plan > $OUT/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
  for k in k2t_thread_walk k2_true_cost k2a_prepare; do
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $OUT/${TAG}_$k -f \
       python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plan > $OUT/${TAG}_ncu_$k.log 2>&1; echo "ncu full $k rc=$?"
  done
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k2t_thread_walk -s 3 -c 1 -o $OUT/${TAG}_k2t_c3 -f \
       python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-plan > $OUT/${TAG}_ncu_k2t_c3.log 2>&1; echo "ncu full k2t c3 rc=$?"
fi
