#!/bin/bash
# GPU call 45: ncu --set full of the chess_4k kernel in the new default shape (640 x 1) and in the old one (256 x 2), same session.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
python tools/jit_variants.py chess_4k "" 1 > gpurun_out/c45_plain_chess.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:maray_jit -s 1 -c 1 -o gpurun_out/c45_chess4k_640 python tools/jit_variants.py chess_4k "" 1 > gpurun_out/c45_ncu_640.log 2>&1
ncu --set full --clock-control none -k regex:maray_jit -s 1 -c 1 -o gpurun_out/c45_chess4k_256 python tools/jit_variants.py chess_4k "MARAY_JIT_BLOCK=256" 1 > gpurun_out/c45_ncu_256.log 2>&1
for n in 640 256; do ncu -i gpurun_out/c45_chess4k_$n.ncu-rep --page raw --csv > gpurun_out/c45_chess4k_${n}_raw.csv 2>/dev/null; done
echo done
