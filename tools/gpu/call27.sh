#!/bin/bash
# GPU call 27: chess_4k, instruction supply: block barriers every N statements x block shapes (no_instruction stalls are 1.0 per issue now).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 900 python tools/jit_variants.py chess_4k ";MARAY_JIT_SYNC_EVERY=128;MARAY_JIT_SYNC_EVERY=512;MARAY_JIT_SYNC_EVERY=2048;MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=1,MARAY_JIT_SYNC_EVERY=256;MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=1,MARAY_JIT_SYNC_EVERY=1024;MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=4;MARAY_JIT_LINEINFO=0;MARAY_JIT_MAXREG=112;MARAY_JIT_MAXREG=104" 5 > gpurun_out/c27_variants_chess4k.jsonl 2> gpurun_out/c27_variants.err
echo done
