#!/bin/bash
# GPU call 60: the chess tests at 1024 x 1024 (now rendered in the 1 024 x 1 shape), bands and the auto back end.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
( time timeout 70 python -m pytest tests -m gpu -q -x -k "chess_against_oracle_golden or backends_agree or ragged or gen_to_image or pipelined or ipc" 2>&1 | tail -6 ) > gpurun_out/c60_pytest.log 2>&1
echo done
