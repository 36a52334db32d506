#!/bin/bash
# GPU call 12: step(sin(u)) as the sign of the sine (exact rewrite): chess kernels with and without, the same in exact
# libm mode, what the out-of-range branches cost; parity tests that render chess.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 600 python tools/jit_variants.py chess_4k ";MARAY_JIT_SIGN_OF_SINE=0;MARAY_JIT_NOSLOW=1;MARAY_LIBM=glibc" 5 > gpurun_out/c12_variants_chess4k.jsonl 2> gpurun_out/c12_variants.err
timeout 600 python tools/jit_variants.py chess_1k ";MARAY_JIT_SIGN_OF_SINE=0;MARAY_JIT_NOSLOW=1" 5 > gpurun_out/c12_variants_chess1k.jsonl 2>> gpurun_out/c12_variants.err
( time timeout 1200 python -m pytest tests -m gpu -q -x -k "chess or backends_agree or known_answers or nan" 2>&1 | tail -15 ) > gpurun_out/c12_pytest.log 2>&1
echo done
