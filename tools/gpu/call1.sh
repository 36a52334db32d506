#!/bin/bash
# GPU call 1 of round 2: the full -m gpu suite (new full-size parity tests, the three optional forms),
# the bench line with the parity key and pageable e2e, and three experiments.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MARAY_JIT_CACHE=$PWD/.jitcache
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
nproc >> gpurun_out/c1_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | tail -80 ) > gpurun_out/c1_pytest.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
MARAY_PIPELINE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_pipeline.json 2> gpurun_out/c1_bench_pipeline.err
timeout 300 python tools/jit_variants.py chess_4k ";MARAY_SCHEDULE=2;MARAY_SCHEDULE=1;MARAY_SCHEDULE=3" 5 > gpurun_out/c1_variants_schedule.jsonl 2>&1
for uni in 0 1; do
  MARAY_INTERP_UNIFORM=$uni timeout 300 python bench.py --workload chess_1k --backend interp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c1_interp_chess1k_uni$uni.json 2> gpurun_out/c1_interp_uni$uni.err
done
MARAY_INTERP_UNIFORM=1 timeout 300 python bench.py --workload sdf --size 2048x1080 --backend interp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c1_interp_sdf_uni1.json 2>> gpurun_out/c1_interp_uni1.err
MARAY_INTERP_UNIFORM=0 timeout 300 python bench.py --workload sdf --size 2048x1080 --backend interp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c1_interp_sdf_uni0.json 2>> gpurun_out/c1_interp_uni0.err
echo done
