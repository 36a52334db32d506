#!/bin/bash
# GPU call 9: full GPU suite after the neg-riding fix, interpreter sweeps, ncu of a deep chain kernel and of the interpreter.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | tail -70 ) > gpurun_out/c9_pytest.log 2>&1
timeout 300 python tools/interp_sweep.py chess_1k ";64,2;128,1" > gpurun_out/c9_sweep_chess1k.jsonl 2> gpurun_out/c9_sweep.err
timeout 300 python tools/interp_sweep.py chess_4k ";128,2;64,2;128,1" 2 > gpurun_out/c9_sweep_chess4k.jsonl 2>> gpurun_out/c9_sweep.err
timeout 300 python tools/interp_sweep.py sdf ";256,4;128,4" > gpurun_out/c9_sweep_sdf.jsonl 2>> gpurun_out/c9_sweep.err
timeout 300 python tools/interp_sweep.py textured ";128,2" > gpurun_out/c9_sweep_textured.jsonl 2>> gpurun_out/c9_sweep.err
MARAY_DEEP_VALUES=20000 timeout 600 python tools/interp_sweep.py deep:1024x512 "" 2 > gpurun_out/c9_sweep_deep20k.jsonl 2>> gpurun_out/c9_sweep.err
MARAY_DEEP_VALUES=20000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:maray_jit -s 4 -c 2 -o gpurun_out/c9_deep20k_chain python tools/jit_variants.py deep:1024x1024 "" 1 > gpurun_out/c9_ncu_deep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:maray_interp -c 1 -o gpurun_out/c9_interp_chess4k python bench.py --workload chess_4k --backend interp --steps 1 --warmup 1 --no-cpu-baseline --configs none --no-first-frame > gpurun_out/c9_ncu_interp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:maray_interp -c 1 -o gpurun_out/c9_interp_sdf python bench.py --workload sdf --backend interp --steps 1 --warmup 1 --no-cpu-baseline --configs none --no-first-frame > gpurun_out/c9_ncu_interp_sdf.log 2>&1
echo done
