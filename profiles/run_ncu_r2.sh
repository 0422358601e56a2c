#!/bin/bash
# Round-2 ncu evidence (single GPU, 2^24 particles, steady state): launch list of three pipelined observations and
# full-set captures of the fused instantiate+expansion kernel and of the survivor scan.
# Usage (under gpurun): bash profiles/run_ncu_r2.sh <tag>
TAG=${1:-r2}
CMD="python bench.py --steps 2 --warmup 3 --burnin 96 --no-cpu"
OUT=gpurun_out
mkdir -p $OUT
# 17 launches per pipelined observation: classify + solve (x3 queued rounds, two of them no-ops), tile_sums (no-op),
# tile_scan1, tile_prefix, tooth_bases, scan_apply, steplog, prestep, count_next, expand<1,INST> (sample), bracket, expand<2,INST>
$CMD > $OUT/ncu_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1530 -c 51 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
for K in ${KERNELS:-k_expandILi2ELi2ELb1 k_scan_apply}; do
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$K -s 90 -c 1 -f -o $OUT/prof_${TAG}_$K $CMD > $OUT/ncu_full_${TAG}_$K.log 2>&1
done
ls -la $OUT | grep $TAG
