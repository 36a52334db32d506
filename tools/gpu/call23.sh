#!/bin/bash
# GPU call 23: which knob makes the NVRTC-12.8 build of the 7-kernel chain (segment size 3072) render the right frame?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_DEEP_VALUES=20000
S="MARAY_JIT_CHAIN_SEGMENT_VALUES=3072"
python tools/jit_variants.py deep:1024x1024 ";$S;$S,MARAY_LIBM_EXPLOG=poly;$S,MARAY_LIBM=cuda;$S,MARAY_LIBM=glibc;$S,MARAY_JIT_SCRATCH=0;$S,MARAY_JIT_BATCH_WIDTH=4;$S,MARAY_JIT_LINEINFO=0;$S,MARAY_JIT_CONST_BANK=0;MARAY_JIT_CHAIN_SEGMENT_VALUES=3000;MARAY_JIT_CHAIN_SEGMENT_VALUES=3200;MARAY_JIT_CHAIN_SEGMENT_VALUES=4096" 1 2>&1 | grep -o '"variant": "[^"]*"\|"rgb_sha": "[^"]*"\|"segments": [0-9]*' | paste - - - > gpurun_out/c23.log 2>&1
echo done
