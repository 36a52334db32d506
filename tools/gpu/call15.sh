#!/bin/bash
# GPU call 15: deep scene (20 000 values): smaller chain segments x fewer registers (more resident warps).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
V=""
for seg in 6144 3072 1536 768; do for mb in 2 3 4; do V="$V;MARAY_JIT_CHAIN_SEGMENT_VALUES=$seg,MARAY_JIT_MIN_BLOCKS=$mb"; done; done
MARAY_DEEP_VALUES=20000 timeout 900 python tools/jit_variants.py deep:1024x1024 "${V:1}" 3 > gpurun_out/c15_variants_deep20k.jsonl 2> gpurun_out/c15_variants.err
echo done
