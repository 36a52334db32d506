#!/bin/bash
# GPU call 13: glibc's exp/log as the default exp/log (deep scene A/B against the polynomial versions and the exact mode),
# chess with step(v + c) folded into comparisons and the sign of the sine from a reduction by pi; parity subset.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
MARAY_DEEP_VALUES=20000 timeout 600 python tools/jit_variants.py deep:1024x1024 ";MARAY_LIBM_EXPLOG=poly;MARAY_LIBM=glibc;MARAY_JIT_BATCH_WIDTH=4" 3 > gpurun_out/c13_variants_deep20k.jsonl 2> gpurun_out/c13_variants.err
timeout 600 python tools/jit_variants.py chess_4k ";MARAY_JIT_SIGN_OF_SINE=0" 5 > gpurun_out/c13_variants_chess4k.jsonl 2>> gpurun_out/c13_variants.err
timeout 300 python tools/interp_sweep.py "deep:1024x512" ";" > gpurun_out/c13_interp_deep.jsonl 2>> gpurun_out/c13_variants.err
( time timeout 1500 python -m pytest tests -m gpu -q -x -k "deep or glibc or chess or backends_agree or batched or nan or known" 2>&1 | tail -15 ) > gpurun_out/c13_pytest.log 2>&1
echo done
