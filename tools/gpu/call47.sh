#!/bin/bash
# GPU call 47: persistent form of the straight-line kernel (one resident block walks the band) against one launch per block.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 300 python tools/jit_variants.py chess_4k ";MARAY_JIT_PERSISTENT=1;MARAY_JIT_PERSISTENT=1,MARAY_JIT_BLOCK=768,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_PERSISTENT=1,MARAY_JIT_BLOCK=1024,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_PERSISTENT=1,MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=2" 5 > gpurun_out/c47_variants_chess4k.jsonl 2> gpurun_out/c47.err
timeout 300 python tools/jit_variants.py chess_1k ";MARAY_JIT_PERSISTENT=1" 5 > gpurun_out/c47_variants_chess1k.jsonl 2>> gpurun_out/c47.err
echo done
