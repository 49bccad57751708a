#!/bin/bash
# r02an: compute-sanitizer over every entry point with the final binary (the goal-set plans now go through
# omgb_chomp_plan_goalset and the two-step learner kernel)
TAG=r02an
mkdir -p gpurun_out
timeout 110 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/memcheck_smoke_$TAG.log 2>&1; grep -c "plan ok" gpurun_out/memcheck_smoke_$TAG.log; tail -2 gpurun_out/memcheck_smoke_$TAG.log
timeout 110 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/racecheck_smoke_$TAG.log 2>&1; grep -c "plan ok" gpurun_out/racecheck_smoke_$TAG.log; tail -2 gpurun_out/racecheck_smoke_$TAG.log
