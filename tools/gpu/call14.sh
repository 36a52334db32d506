#!/bin/bash
# GPU call 14: ncu --set full of two chain kernels of the deep scene (20 000 values) and of the chess_4k kernel (new rewrites).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_DEEP_VALUES=20000
python tools/jit_variants.py deep:1024x1024 "" 1 > gpurun_out/c14_plain_deep.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:maray_jit -s 1 -c 2 -o gpurun_out/c14_deep20k python tools/jit_variants.py deep:1024x1024 "" 1 > gpurun_out/c14_ncu_deep.log 2>&1
python tools/jit_variants.py chess_4k "" 1 > gpurun_out/c14_plain_chess.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:maray_jit -s 1 -c 1 -o gpurun_out/c14_chess4k python tools/jit_variants.py chess_4k "" 1 > gpurun_out/c14_ncu_chess.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
