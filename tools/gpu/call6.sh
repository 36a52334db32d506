#!/bin/bash
# GPU call 6: AUTO back end (first frame), deep helper register experiments, full GPU suite.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | tail -70 ) > gpurun_out/c6_pytest.log 2>&1
timeout 600 python tools/first_frame.py chess_4k > gpurun_out/c6_first_frame.jsonl 2> gpurun_out/c6_first_frame.err
timeout 600 python tools/first_frame.py chess_1k >> gpurun_out/c6_first_frame.jsonl 2>> gpurun_out/c6_first_frame.err
timeout 600 python tools/first_frame.py sdf >> gpurun_out/c6_first_frame.jsonl 2>> gpurun_out/c6_first_frame.err
timeout 900 python tools/first_frame.py deep auto,nvrtc >> gpurun_out/c6_first_frame.jsonl 2>> gpurun_out/c6_first_frame.err
MARAY_DEEP_VALUES=20000 timeout 900 python tools/jit_variants.py deep:1024x1024 ";MARAY_JIT_BLOCK=192;MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=3;MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=2;MARAY_JIT_BLOCK=192,MARAY_JIT_BATCH_WIDTH=4;MARAY_JIT_BLOCK=128,MARAY_JIT_MIN_BLOCKS=2,MARAY_JIT_BATCH_WIDTH=4;MARAY_JIT_BLOCK=320,MARAY_JIT_MIN_BLOCKS=1;MARAY_JIT_FRAME_MB=8192" 3 > gpurun_out/c6_variants_deep20k.jsonl 2> gpurun_out/c6_variants.err
echo done
