#!/bin/bash
# ncu evidence for one round: launch list (per-launch device time) and a full-set capture of
# the four data-path kernels at steady state (2^24 particles, after the burn-in).
# Usage (under gpurun): bash profiles/run_ncu.sh <tag>
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 1 --burnin 48 --no-cpu"
OUT=gpurun_out
mkdir -p $OUT
$CMD > $OUT/ncu_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 392 -c 40 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
for K in k_expand k_instantiate k_select k_scan_apply k_tile_sums; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 50 -c 1 -f -o $OUT/prof_${TAG}_$K $CMD > $OUT/ncu_full_${TAG}_$K.log 2>&1
done
ls -la $OUT
