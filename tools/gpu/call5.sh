#!/bin/bash
# GPU call 5: the chain form of large programs (parity + timing against the one-unit form), interpreter dispatch A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 1200 python -m pytest tests -m gpu -q -x -k "batched or deep or smoke" 2>&1 | tail -40 ) > gpurun_out/c5_pytest.log 2>&1
MARAY_DEEP_VALUES=20000 timeout 900 python tools/jit_variants.py deep:1024x1024 ";MARAY_JIT_CHAIN=0;MARAY_JIT_SEGMENT_VALUES=3072;MARAY_JIT_SEGMENT_VALUES=10000;MARAY_JIT_FRAME_MB=64" 3 > gpurun_out/c5_variants_deep20k.jsonl 2> gpurun_out/c5_variants.err
timeout 900 python tools/jit_variants.py deep ";MARAY_JIT_CHAIN=0" 2 > gpurun_out/c5_variants_deep_full.jsonl 2>> gpurun_out/c5_variants.err
for d in shared private; do
  MARAY_INTERP_DISPATCH=$d timeout 300 python tools/interp_sweep.py chess_1k ";64,2;128,1" > gpurun_out/c5_sweep_chess1k_$d.jsonl 2>> gpurun_out/c5_sweep.err
  MARAY_INTERP_DISPATCH=$d timeout 300 python tools/interp_sweep.py chess_4k "" 2 > gpurun_out/c5_sweep_chess4k_$d.jsonl 2>> gpurun_out/c5_sweep.err
  MARAY_INTERP_DISPATCH=$d timeout 300 python tools/interp_sweep.py sdf ";256,4" > gpurun_out/c5_sweep_sdf_$d.jsonl 2>> gpurun_out/c5_sweep.err
done
echo done
