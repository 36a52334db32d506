#!/bin/bash
# GPU call 42: chess_4k: one block per SM, block size sweep (registers follow from the block size).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 300 python tools/jit_variants.py chess_4k ";MARAY_JIT_BLOCK=768,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=640,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=896,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=2;MARAY_JIT_BLOCK=384,MARAY_JIT_MIN_BLOCKS=2;MARAY_JIT_BLOCK=576,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=704,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=832,MARAY_JIT_MIN_BLOCKS=1" 5 > gpurun_out/c42_variants_chess4k.jsonl 2> gpurun_out/c42.err
echo done
