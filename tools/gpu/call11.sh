#!/bin/bash
# GPU call 11: exact libm mode on the device (parity + cost), out-of-line sin on chess, band drain per launch shape.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 900 python -m pytest tests/test_glibc_libm.py -m gpu -q -rA 2>&1 | tail -40 ) > gpurun_out/c11_pytest.log 2>&1
timeout 600 python tools/jit_variants.py chess_4k ";MARAY_LIBM=glibc;MARAY_JIT_INLINE_TRANS_BELOW=0;MARAY_LIBM=cuda" 5 > gpurun_out/c11_variants_chess4k.jsonl 2> gpurun_out/c11_variants.err
MARAY_DEEP_VALUES=20000 timeout 600 python tools/jit_variants.py deep:1024x1024 ";MARAY_LIBM=glibc" 3 > gpurun_out/c11_variants_deep20k.jsonl 2>> gpurun_out/c11_variants.err
timeout 600 python tools/band_tail.py chess_4k ";MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=4;MARAY_JIT_BLOCK=64,MARAY_JIT_MIN_BLOCKS=8;MARAY_JIT_BLOCK=512,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=192,MARAY_JIT_MIN_BLOCKS=2" 8 20 > gpurun_out/c11_band_tail.jsonl 2> gpurun_out/c11_band_tail.err
for s in "" "MARAY_LIBM=glibc"; do env $s timeout 300 python tools/interp_sweep.py chess_1k ";" >> gpurun_out/c11_interp.jsonl 2>> gpurun_out/c11_interp.err; done
echo done
