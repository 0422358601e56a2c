#!/bin/bash
# compute-sanitizer over the parity suite (small cases: the tools slow every kernel down 10-50x).
# Usage (under gpurun): bash profiles/run_sanitizer.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
SMALL='not 2_pow and not at_2_pow and not large and not 1_pow and not one_call and not properties'
# memcheck: out-of-bounds / misaligned accesses, leaks of device allocations at exit
timeout 1500 compute-sanitizer --tool memcheck --leak-check full --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_statepriors.py tests/test_gpu_summaries.py tests/test_gpu_f32.py tests/test_gpu_sharded.py \
    -q -x -m gpu -k "$SMALL" > $OUT/sanitizer_memcheck_$TAG.log 2>&1
echo "memcheck exit code $?" >> $OUT/sanitizer_memcheck_$TAG.log
# racecheck: shared-memory hazards inside the kernels (block scans, candidate staging, walk queue, histograms)
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_summaries.py -q -x -m gpu -k "$SMALL and not drift and not long_stream" > $OUT/sanitizer_racecheck_$TAG.log 2>&1
echo "racecheck exit code $?" >> $OUT/sanitizer_racecheck_$TAG.log
tail -5 $OUT/sanitizer_memcheck_$TAG.log $OUT/sanitizer_racecheck_$TAG.log
