#!/bin/bash
# GPU call 57: whole GPU suite, bench line, smoke and launch list: final build of the round (after the helper-table change).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 ) > gpurun_out/c57_pytest.log 2>&1
( time python bench.py --steps 10 --warmup 3 > gpurun_out/c57_bench_n1.json 2> gpurun_out/c57_bench_n1.err ) 2> gpurun_out/c57_bench_n1.time
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c57_smoke.log 2>&1
S="--steps 2 --warmup 1 --no-cpu-baseline --configs none --no-first-frame"
python bench.py $S > gpurun_out/c57_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c57_launches.csv python bench.py $S > gpurun_out/c57_ncu_launch.log 2>&1
echo done
