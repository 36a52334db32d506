#!/bin/bash
# GPU call 33 (8 GPUs): the bench line at N = 8, 4 and 2 with the round's final build; the multi-GPU parity tests.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
nvidia-smi -L > gpurun_out/c33_gpus.txt
( timeout 300 python -m pytest tests -m gpu -q -x -k "multi_gpu or ipc" 2>&1 | tail -5 ) > gpurun_out/c33_pytest.log 2>&1
for n in 8 4 2; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/c33_bench_n$n.json 2> gpurun_out/c33_bench_n$n.err ) 2> gpurun_out/c33_bench_n$n.time
done
echo done
