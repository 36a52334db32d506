#!/bin/bash
# GPU call 48 (2 GPUs): completion counters over NVLink against the NCCL reduce at N = 2; the IPC + counters test; in-process 2-GPU test.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( timeout 600 python -m pytest tests -m gpu -q -x -k "ipc or multi_gpu" 2>&1 | tail -6 ) > gpurun_out/c48_pytest.log 2>&1
S="--steps 20 --warmup 5 --no-cpu-baseline --configs none --no-first-frame"
for sig in counters nccl; do
  ( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29841 bench.py --gpus 2 $S --completion $sig > gpurun_out/c48_bench_n2_$sig.json 2> gpurun_out/c48_bench_n2_$sig.err ) 2> gpurun_out/c48_bench_n2_$sig.time
done
echo done
