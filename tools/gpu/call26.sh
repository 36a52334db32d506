#!/bin/bash
# GPU call 26: reference arm; launch list of a short bench run; ncu --set full of the final chess_4k kernel; GPU tests of the new cases.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c26_bench_n1_reference.json 2> gpurun_out/c26_ref.err ) 2> gpurun_out/c26_ref.time
( timeout 900 python -m pytest tests -m gpu -q -x -k "sign_only or segmentations or glibc" 2>&1 | tail -6 ) > gpurun_out/c26_pytest.log 2>&1
S="--steps 2 --warmup 1 --no-cpu-baseline --configs none --no-first-frame"
python bench.py $S > gpurun_out/c26_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c26_launches.csv python bench.py $S > gpurun_out/c26_ncu_launch.log 2>&1
python tools/jit_variants.py chess_4k "" 1 > gpurun_out/c26_plain_chess.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:maray_jit -s 1 -c 1 -o gpurun_out/c26_chess4k python tools/jit_variants.py chess_4k "" 1 > gpurun_out/c26_ncu_chess.log 2>&1
echo done
