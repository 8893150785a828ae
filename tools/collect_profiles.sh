#!/bin/bash
# After tools/gpu_round2.sh <tag> has left its captures in gpurun_out/: the committed summaries under profiles/.
# Usage (repo root, no GPU needed): bash tools/collect_profiles.sh <tag>
set -u
TAG=${1:-r02}
N=1048576
python tools/ncu_metrics.py $TAG $N
for wl in c5 c2; do
  cp gpurun_out/${TAG}_launches_${wl}.csv profiles/${TAG}_launches_${wl}.csv
  for k in k2a_prepare k2t_thread_walk k2_true_cost; do
    python tools/ncu_summary.py gpurun_out/${TAG}_${k}_${wl}.ncu-rep $N profiles/${TAG}_${k}_${wl}_ncu_summary.txt > /dev/null
  done
  python tools/sass_hotspots.py gpurun_out/${TAG}_k2_true_cost_${wl}.ncu-rep k2_true_costILi4 > profiles/${TAG}_k2_true_cost_${wl}_hotspots.txt 2>&1
  python tools/sass_hotspots.py gpurun_out/${TAG}_k2t_thread_walk_${wl}.ncu-rep k2t_thread_walk > profiles/${TAG}_k2t_thread_walk_${wl}_hotspots.txt 2>&1
done
cp gpurun_out/${TAG}_gpu.txt profiles/${TAG}_gpu.txt 2>/dev/null
ls -la profiles | grep ${TAG}_
