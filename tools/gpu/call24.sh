#!/bin/bash
# GPU call 24: everything again with the statically linked NVRTC 12.9: the whole GPU suite, chess and deep kernels, bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 300 python tools/jit_variants.py chess_4k ";MARAY_JIT_SIGN_OF_SINE=0" 5 > gpurun_out/c24_variants_chess4k.jsonl 2> gpurun_out/c24_variants.err
MARAY_DEEP_VALUES=20000 timeout 300 python tools/jit_variants.py deep:1024x1024 ";MARAY_JIT_CHAIN_SEGMENT_VALUES=3072;MARAY_LIBM=glibc" 3 > gpurun_out/c24_variants_deep20k.jsonl 2>> gpurun_out/c24_variants.err
( time timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/c24_pytest.log 2>&1
echo done
