#!/bin/bash
# One GPU-box visit for the committed profiles of round 2: launch lists and `ncu --set full` captures of the kernels of one
# step of the default bench command (C5) and of C2; the N2 A/B captures of the thread walker.  Run only after the same
# commands have exited 0 without ncu.  Usage (from the repo root, via gpurun): bash tools/gpu_round2.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plan --no-extra"
for wl in c5 c2; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches_${wl}.csv \
     $B --workload $wl > $OUT/${TAG}_ncu_list_${wl}.log 2>&1; echo "ncu list $wl rc=$?"
  for k in k2a_prepare k2t_thread_walk k2_true_cost; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $OUT/${TAG}_${k}_${wl} -f \
       $B --workload $wl > $OUT/${TAG}_ncu_${k}_${wl}.log 2>&1; echo "ncu full $k $wl rc=$?"
  done
done
# (TILEAB_MODES= skips it)
# N2 A/B: the thread walker on the frontier-shaped batch, look-ups from L2 (__ldg) vs the shared-memory tile
for mode in ${TILEAB_MODES-ldg smem_tile}; do
  envs=""; [ $mode = smem_tile ] && envs="PPE_MAP_TILE=1"
  timeout 900 env $envs ncu --set full --clock-control none --import-source on -k regex:k2t_thread_walk -s 3 -c 1 -o $OUT/${TAG}_k2t_tileab_${mode} -f \
     python tools/map_tile_ab.py $mode c4 > $OUT/${TAG}_ncu_tileab_${mode}.log 2>&1; echo "ncu tile A/B $mode rc=$?"
done
ls -la $OUT/*.ncu-rep
