#!/bin/bash
# GPU call 49 (2 GPUs): why the counter signal serialises the bands at N = 2 (per-rank step times).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
export MARAY_BENCH_DEBUG=1
S="--steps 10 --warmup 3 --no-cpu-baseline --configs none --no-first-frame"
for sig in counters nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29842 bench.py --gpus 2 $S --completion $sig > gpurun_out/c49_bench_n2_$sig.json 2> gpurun_out/c49_bench_n2_$sig.err
done
nvidia-smi topo -m > gpurun_out/c49_topo.txt 2>&1
echo done
