#!/bin/bash
# GPU call 8 (2 GPUs): the N=2 bench line (IPC peer stores, shared pinned host frame, in-process path) and the
# in-process multi-GPU parity test; interpreter sweep with the neg-riding handlers.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
nvidia-smi -L > gpurun_out/c8_gpus.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c8_bench_n2.json 2> gpurun_out/c8_bench_n2.err ) 2> gpurun_out/c8_bench_n2.time
( timeout 600 python -m pytest tests -m gpu -q -x -k "multi_gpu or interp or auto" 2>&1 | tail -15 ) > gpurun_out/c8_pytest.log 2>&1
timeout 300 python tools/interp_sweep.py chess_1k ";64,2;128,1" > gpurun_out/c8_sweep_chess1k.jsonl 2> gpurun_out/c8_sweep.err
timeout 300 python tools/interp_sweep.py chess_4k "" 2 > gpurun_out/c8_sweep_chess4k.jsonl 2>> gpurun_out/c8_sweep.err
echo done
