#!/bin/bash
# GPU call 52 (2 GPUs): per-step alignment by a stream-ordered all-reduce instead of a host-blocking barrier.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_BENCH_DEBUG=1
S="--steps 20 --warmup 5 --no-cpu-baseline --configs none --no-first-frame"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus 2 $S > gpurun_out/c52_bench_n2.json 2> gpurun_out/c52_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29872 bench.py --gpus 2 $S --completion counters > gpurun_out/c52_bench_n2_counters.json 2> gpurun_out/c52_bench_n2_counters.err
echo done
