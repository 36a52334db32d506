#!/bin/bash
# GPU call 40: chess_4k, final-build kernel: more resident warps through lower register caps (128-thread blocks x 4..8 per SM).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
V=";MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=5;MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=6;MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=7;MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=8"
timeout 300 python tools/jit_variants.py chess_4k "$V" 5 > gpurun_out/c40_variants_chess4k.jsonl 2> gpurun_out/c40.err
echo done
