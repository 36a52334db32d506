#!/bin/bash
# GPU call 62: smoke() of the final library build (the last seconds of the round's GPU budget).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
timeout 22 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c62_smoke.log 2>&1
echo done
