#!/bin/bash
# GPU call 4: interpreter with the 13-instruction asm loop and the new shape heuristic; full GPU suite.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | tail -60 ) > gpurun_out/c4_pytest.log 2>&1
timeout 300 python tools/interp_sweep.py chess_1k ";64,2;128,1;256,2;416,2;384,2;192,4;512,1;256,1" > gpurun_out/c4_sweep_chess1k.jsonl 2> gpurun_out/c4_sweep.err
timeout 300 python tools/interp_sweep.py chess_4k ";384,2;192,2" 2 > gpurun_out/c4_sweep_chess4k.jsonl 2>> gpurun_out/c4_sweep.err
timeout 300 python tools/interp_sweep.py sdf ";128,4;256,4;480,4;480,2;384,2;192,2;960,2;320,2" > gpurun_out/c4_sweep_sdf.jsonl 2>> gpurun_out/c4_sweep.err
MARAY_DEEP_VALUES=20000 timeout 600 python tools/interp_sweep.py deep:1024x512 ";64,1;160,1;64,2" 2 > gpurun_out/c4_sweep_deep20k.jsonl 2>> gpurun_out/c4_sweep.err
timeout 300 python tools/interp_sweep.py textured ";128,2;256,2;512,2;256,4;384,2" > gpurun_out/c4_sweep_textured.jsonl 2>> gpurun_out/c4_sweep.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:maray_interp -c 1 -o gpurun_out/c4_interp_chess1k python bench.py --workload chess_1k --backend interp --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c4_ncu.log 2>&1
echo done
