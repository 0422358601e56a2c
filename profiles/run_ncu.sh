#!/bin/bash
# ncu evidence for one round: launch list (per-launch device time) and a full-set capture of
# the data-path kernels at steady state (2^24 particles, after the burn-in).
# Usage (under gpurun): bash profiles/run_ncu.sh <tag>
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 1 --burnin 48 --no-cpu"
OUT=gpurun_out
mkdir -p $OUT
# 11 kernels per observation: prestep, expand<1> (sample), select (bracket), expand<2> (fused), select (solve),
# tile_fixup, tile_sums (skipped when the fused records are usable), tile_prefix, scan_apply, instantiate, steplog
$CMD > $OUT/ncu_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 539 -c 44 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
for K in ${KERNELS:-k_expandILi2ELi2 k_instantiate k_scan_apply k_select}; do
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$K -s 50 -c 1 -f -o $OUT/prof_${TAG}_$K $CMD > $OUT/ncu_full_${TAG}_$K.log 2>&1
done
ls -la $OUT | grep $TAG
