#!/bin/bash
# GPU call 3: interpreter with the inline-PTX jump-table inner loop -- parity tests, shape sweep, tree-vs-table A/B, one ncu capture.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 900 python -m pytest tests -m gpu -q -x -k "interp or agree or ragged or gen_to_image or chess_4k" 2>&1 | tail -40 ) > gpurun_out/c3_pytest.log 2>&1
timeout 300 python tools/interp_sweep.py chess_1k > gpurun_out/c3_sweep_chess1k.jsonl 2> gpurun_out/c3_sweep.err
timeout 300 python tools/interp_sweep.py sdf > gpurun_out/c3_sweep_sdf.jsonl 2>> gpurun_out/c3_sweep.err
MARAY_DEEP_VALUES=20000 timeout 600 python tools/interp_sweep.py deep:1024x512 ";64,1;128,1;256,1;64,2;128,2" 2 > gpurun_out/c3_sweep_deep20k.jsonl 2>> gpurun_out/c3_sweep.err
timeout 300 python tools/interp_sweep.py textured ";128,1;128,2;128,4;256,4" > gpurun_out/c3_sweep_textured.jsonl 2>> gpurun_out/c3_sweep.err
MARAY_INTERP_DISPATCH=tree timeout 300 python tools/interp_sweep.py chess_1k ";128,2" > gpurun_out/c3_sweep_chess1k_tree.jsonl 2>> gpurun_out/c3_sweep.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:maray_interp -c 1 -o gpurun_out/c3_interp_chess1k python bench.py --workload chess_1k --backend interp --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c3_ncu.log 2>&1
echo done
