#!/bin/bash
# GPU call 35 (8 GPUs): bench at N = 8 and 4 with the reduce-to-root completion signal and the overlapped e2e halves.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
for n in 8 4; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2981$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/c35_bench_n$n.json 2> gpurun_out/c35_bench_n$n.err ) 2> gpurun_out/c35_bench_n$n.time
done
echo done
