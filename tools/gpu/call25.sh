#!/bin/bash
# GPU call 25: the bench line (N = 1) with the round's rewrites and the static NVRTC; launch list of the same command.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time python bench.py --steps 10 --warmup 3 > gpurun_out/c25_bench_n1.json 2> gpurun_out/c25_bench_n1.err ) 2> gpurun_out/c25_bench_n1.time
tail -c 600 gpurun_out/c25_bench_n1.err
echo done
