#!/bin/bash
# GPU call 61: bench line of chess_1k alone (config 1) with the shape chosen from the frame size.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 45 python bench.py --steps 10 --warmup 3 --workload chess_1k --no-cpu-baseline --configs none --no-first-frame > gpurun_out/c61_bench_chess1k.json 2> gpurun_out/c61.err
echo done
