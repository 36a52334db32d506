#!/bin/bash
# GPU call 41: chess_4k: one large block per SM (24 / 32 warps at 80 / 64 registers) with block barriers every N statements.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 300 python tools/jit_variants.py chess_4k ";MARAY_JIT_BLOCK=1024,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=1024,MARAY_JIT_MIN_BLOCKS=1,MARAY_JIT_SYNC_EVERY=128;MARAY_JIT_BLOCK=1024,MARAY_JIT_MIN_BLOCKS=1,MARAY_JIT_SYNC_EVERY=512;MARAY_JIT_BLOCK=768,MARAY_JIT_MIN_BLOCKS=1,MARAY_JIT_SYNC_EVERY=128;MARAY_JIT_BLOCK=768,MARAY_JIT_MIN_BLOCKS=1,MARAY_JIT_SYNC_EVERY=512;MARAY_JIT_BLOCK=768,MARAY_JIT_MIN_BLOCKS=1" 5 > gpurun_out/c41_variants_chess4k.jsonl 2> gpurun_out/c41.err
echo done
