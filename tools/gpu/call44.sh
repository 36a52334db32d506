#!/bin/bash
# GPU call 44: constants in use order (one table entry per constant operand, neighbours share a 128-bit LDCU): chess_4k at several shapes, deep at 20 000 values.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 300 python tools/jit_variants.py chess_4k ";MARAY_JIT_CONST_ORDER=1;MARAY_JIT_CONST_ORDER=1,MARAY_JIT_BLOCK=768,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_CONST_ORDER=1,MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_CONST_ORDER=1,MARAY_JIT_BLOCK=1024,MARAY_JIT_MIN_BLOCKS=1" 5 > gpurun_out/c44_variants_chess4k.jsonl 2> gpurun_out/c44.err
MARAY_DEEP_VALUES=20000 timeout 300 python tools/jit_variants.py deep:1024x1024 ";MARAY_JIT_CONST_ORDER=1" 5 > gpurun_out/c44_variants_deep20k.jsonl 2>> gpurun_out/c44.err
echo done
