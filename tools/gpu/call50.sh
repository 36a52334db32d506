#!/bin/bash
# GPU call 50 (2 GPUs): call 48's configuration again with per-rank step times (counters, nccl, counters).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_BENCH_DEBUG=1
S="--steps 20 --warmup 5 --no-cpu-baseline --configs none --no-first-frame"
i=0
for sig in counters nccl counters; do
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2985$i bench.py --gpus 2 $S --completion $sig > gpurun_out/c50_bench_n2_${i}_$sig.json 2> gpurun_out/c50_bench_n2_${i}_$sig.err
done
echo done
