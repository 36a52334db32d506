#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_DEEP_VALUES=20000
A="MARAY_JIT_CHAIN_SEGMENT_VALUES=6144,MARAY_JIT_MIN_BLOCKS=2"; B="MARAY_JIT_CHAIN_SEGMENT_VALUES=6144,MARAY_JIT_MIN_BLOCKS=3"; C="MARAY_JIT_CHAIN_SEGMENT_VALUES=6144,MARAY_JIT_MIN_BLOCKS=4"; D="MARAY_JIT_CHAIN_SEGMENT_VALUES=3072,MARAY_JIT_MIN_BLOCKS=2"
for seq in "$D" "$A;$D" "$B;$D" "$C;$D" "$D;$D" "$A;$B;$D" ";$D" "MARAY_JIT_CHAIN_SEGMENT_VALUES=3072;$D"; do
  echo "== $seq"
  python tools/jit_variants.py deep:1024x1024 "$seq" 1 2>&1 | grep -o '"variant": "[^"]*"\|"rgb_sha": "[^"]*"' | paste - -
done > gpurun_out/c22.log 2>&1
echo done
