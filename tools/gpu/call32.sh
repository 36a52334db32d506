#!/bin/bash
# GPU call 32: whole GPU suite and the bench line with the new sine / four-wide helpers (final build of the round).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 ) > gpurun_out/c32_pytest.log 2>&1
( time python bench.py --steps 10 --warmup 3 > gpurun_out/c32_bench_n1.json 2> gpurun_out/c32_bench_n1.err ) 2> gpurun_out/c32_bench_n1.time
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c32_smoke.log 2>&1
echo done
