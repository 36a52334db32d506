#!/bin/bash
# GPU call 46: whole GPU suite, bench line and smoke with the automatic launch shape.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 ) > gpurun_out/c46_pytest.log 2>&1
( time python bench.py --steps 10 --warmup 3 > gpurun_out/c46_bench_n1.json 2> gpurun_out/c46_bench_n1.err ) 2> gpurun_out/c46_bench_n1.time
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c46_smoke.log 2>&1
echo done
