#!/bin/bash
# GPU call 38: deep scene (20 000 values): shared-memory carve-out (computed / driver's / maximum) and launch shapes under the computed carve-out.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_DEEP_VALUES=20000
V=";MARAY_JIT_CARVEOUT=-2;MARAY_JIT_CARVEOUT=100;MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=4;MARAY_JIT_BLOCK=192,MARAY_JIT_MIN_BLOCKS=2;MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=3"
timeout 300 python tools/jit_variants.py deep:1024x1024 "$V" 5 > gpurun_out/c38_variants_deep20k.jsonl 2> gpurun_out/c38.err
unset MARAY_DEEP_VALUES
timeout 200 python tools/jit_variants.py chess_4k ";MARAY_JIT_CARVEOUT=-2" 5 > gpurun_out/c38_variants_chess4k.jsonl 2>> gpurun_out/c38.err
echo done
