#!/bin/bash
# GPU call 51 (2 GPUs): completion counters, one 128-byte line each; store / atomic exchange signal x acquire load / atomic / volatile poll.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_BENCH_DEBUG=1
S="--steps 20 --warmup 5 --no-cpu-baseline --configs none --no-first-frame"
for mode in 0 1 3 4; do
  MARAY_BAND_SIGNAL_MODE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2986$mode bench.py --gpus 2 $S --completion counters > gpurun_out/c51_bench_n2_mode$mode.json 2> gpurun_out/c51_bench_n2_mode$mode.err
done
echo done
