#!/bin/bash
# GPU call 7: the new bench line (all configs, parity, first frame) at N=1, and the reference arm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err ) 2> gpurun_out/c7_bench.time
timeout 300 python tools/first_frame.py chess_4k auto,nvrtc > gpurun_out/c7_first_frame.jsonl 2> gpurun_out/c7_first_frame.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c7_bench_ref.json 2> gpurun_out/c7_bench_ref.err ) 2> gpurun_out/c7_bench_ref.time
echo done
