#!/bin/bash
# GPU call 43: the automatic launch shape (640 x 1 for large straight-line programs) against 256 x 2 on chess_1k, chess_4k and the chess.rs re-run.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
for s in chess_1k chess_4k chess_dsl; do
timeout 300 python tools/jit_variants.py $s ";MARAY_JIT_BLOCK=256" 5 > gpurun_out/c43_variants_$s.jsonl 2>> gpurun_out/c43.err
done
echo done
