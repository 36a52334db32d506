#!/bin/bash
# GPU call 56: deep scene, 20 000 values: batch helpers without the past-n guard on the range flag, exp/log tables in shared memory (default) against global.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_DEEP_VALUES=20000
timeout 300 python tools/jit_variants.py deep:1024x1024 ";MARAY_JIT_SCRATCH_TABLES=0;MARAY_LIBM=glibc" 5 > gpurun_out/c56_variants_deep20k.jsonl 2> gpurun_out/c56.err
( timeout 600 python -m pytest tests -m gpu -q -x -k "deep or batched or transcend or glibc or segment" 2>&1 | tail -5 ) > gpurun_out/c56_pytest.log 2>&1
echo done
