#!/bin/bash
# GPU call 55: ncu --set full of one chain kernel of the deep scene at FULL size (96 867 values, 8192 x 8192, one frame chunk).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 300 ncu --set full --clock-control none -k regex:maray_jit -s 21 -c 1 -o gpurun_out/c55_deep_full python tools/jit_variants.py deep "" 1 > gpurun_out/c55_ncu_deep.log 2>&1
ncu -i gpurun_out/c55_deep_full.ncu-rep --page raw --csv > gpurun_out/c55_deep_full_raw.csv 2>/dev/null
echo done
