#!/bin/bash
# r02af: compute-sanitizer over the step kernel's data-dependent branches (radix select + member cost from registers,
# segmented winners pass, one-barrier block sums) on the config-2 scene, and over every entry point (sanitize_smoke)
TAG=r02af
mkdir -p gpurun_out
export OMGB_STEP_CONFIG=0
B=192 ITERS=2 timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_step_heavy.py > gpurun_out/racecheck_heavy_$TAG.log 2>&1; tail -6 gpurun_out/racecheck_heavy_$TAG.log
B=192 ITERS=2 TOPK=200 timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_step_heavy.py > gpurun_out/racecheck_heavy_k200_$TAG.log 2>&1; tail -4 gpurun_out/racecheck_heavy_k200_$TAG.log
B=192 ITERS=2 timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_step_heavy.py > gpurun_out/memcheck_heavy_$TAG.log 2>&1; tail -4 gpurun_out/memcheck_heavy_$TAG.log
unset OMGB_STEP_CONFIG
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/racecheck_smoke_$TAG.log 2>&1; tail -4 gpurun_out/racecheck_smoke_$TAG.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/memcheck_smoke_$TAG.log 2>&1; tail -4 gpurun_out/memcheck_smoke_$TAG.log
