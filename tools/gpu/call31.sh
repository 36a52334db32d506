#!/bin/bash
# GPU call 31: ncu --set full of one chain kernel of the deep scene (20 000 values) with the new sine and four-wide helpers.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_DEEP_VALUES=20000
python tools/jit_variants.py deep:1024x1024 "" 1 > gpurun_out/c31_plain_deep.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:maray_jit -s 1 -c 1 -o gpurun_out/c31_deep20k python tools/jit_variants.py deep:1024x1024 "" 1 > gpurun_out/c31_ncu_deep.log 2>&1
echo done
