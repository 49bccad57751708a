#!/bin/bash
TAG=r02h
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_goal_set.py -m gpu -q -s -k large_batch 2>&1 | grep -E "chains|assert|passed|failed" | head
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
timeout 300 python bench.py $Q > gpurun_out/ab3_c2.json 2> gpurun_out/ab3_c2.err
python tools/bench_summary.py c2_fused_winners < gpurun_out/ab3_c2.json
timeout 300 python bench.py $Q --waypoints 60 --objects 20 --grid 256 --steps 10 > gpurun_out/ab3_c4.json 2> gpurun_out/ab3_c4.err
python tools/bench_summary.py c4_fused_winners < gpurun_out/ab3_c4.json
python tools/phase_profile.py > gpurun_out/phase_$TAG.txt 2>&1; tail -28 gpurun_out/phase_$TAG.txt
