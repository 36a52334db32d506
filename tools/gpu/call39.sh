#!/bin/bash
# GPU call 39: chess_4k: block barriers every N statements with 256- and 512-thread blocks (instruction-cache sharing), small blocks.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
V=";MARAY_JIT_SYNC_EVERY=64;MARAY_JIT_SYNC_EVERY=256;MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=1,MARAY_JIT_SYNC_EVERY=64;MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=1,MARAY_JIT_SYNC_EVERY=256;MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=1,MARAY_JIT_SYNC_EVERY=1024;MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=4;MARAY_JIT_BLOCK=64,MARAY_JIT_MIN_BLOCKS=8"
timeout 300 python tools/jit_variants.py chess_4k "$V" 5 > gpurun_out/c39_variants_chess4k.jsonl 2> gpurun_out/c39.err
echo done
