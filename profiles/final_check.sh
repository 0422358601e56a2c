#!/bin/bash
# Round-end verification on one B200: full GPU suite on the default build, the optional paths
# (PF_SCAN_REC, ping-pong expansion variant) through the parity + golden tests, a bench line for
# each combination, and the ncu traffic capture of the dominant kernel at the bench's state.
O=gpurun_out; mkdir -p $O
V=$PWD/particles.jl_b200/_variants
T="tests/test_gpu_parity.py tests/test_golden.py"
echo "== suite (default)";   timeout 280 python -m pytest tests -q -m gpu --timeout 120 2>&1 | tail -3
echo "== PF_SCAN_REC=1";     PF_SCAN_REC=1 timeout 150 python -m pytest $T -q -m gpu --timeout 120 2>&1 | tail -3
echo "== pingpong";          PARTICLES_B200_LIB=$V/lib_pingpong.so PF_SCAN_REC=1 timeout 150 python -m pytest $T -q -m gpu --timeout 120 2>&1 | tail -3
b() { python bench.py --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(j['ms_per_step'],4), {k:round(x['share']*j['ms_per_step'],3) for k,x in j['roofline']['per_kernel'].items()})"; }
b base
PF_SCAN_REC=1 b base+rec
PARTICLES_B200_LIB=$V/lib_pingpong.so b pingpong
PARTICLES_B200_LIB=$V/lib_pingpong.so PF_SCAN_REC=1 b pingpong+rec
PARTICLES_B200_LIB=$V/lib_pingpong_e3.so PF_SCAN_REC=1 b pingpong_e3+rec
ncu --set full --clock-control none --import-source on -k regex:k_instantiate -s 99 -c 1 -f -o $O/prof_r1e_k_instantiate \
    python bench.py --steps 2 --warmup 3 --burnin 96 --no-cpu > $O/ncu_r1e_instantiate.log 2>&1
tail -c 600 $O/ncu_r1e_instantiate.log
