#!/bin/bash
# GPU call 59: chess_1k with the shape chosen from the declared frame size (1 024 x 1: 6.92 rounds) against 640 x 1 (11.07 rounds); the launch-shape test.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 200 python tools/jit_variants.py chess_1k ";MARAY_JIT_BLOCK=640,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_BLOCK=768,MARAY_JIT_MIN_BLOCKS=1" 7 > gpurun_out/c59_variants_chess1k.jsonl 2> gpurun_out/c59.err
( timeout 200 python -m pytest tests -m gpu -q -x -k "launch_shapes" 2>&1 | tail -5 ) > gpurun_out/c59_pytest.log 2>&1
echo done
