#!/bin/bash
# GPU call 53 (8 GPUs): chess_4k at N = 8 with the 640 x 1 launch shape and the stream-ordered step alignment (headline only).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_BENCH_DEBUG=1
S="--steps 20 --warmup 5 --no-cpu-baseline --configs none --no-first-frame"
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29881 bench.py --gpus 8 $S > gpurun_out/c53_bench_n8.json 2> gpurun_out/c53_bench_n8.err ) 2> gpurun_out/c53_bench_n8.time
echo done
