#!/bin/bash
# GPU call 58: the new GPU tests (launch shapes of a straight-line program; the added code-generation knobs).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 300 python -m pytest tests -m gpu -q -x -k "launch_shapes or code_generation_knobs" 2>&1 | tail -15 ) > gpurun_out/c58_pytest.log 2>&1
echo done
