#!/bin/bash
TAG=r02c
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs_fullsize.py tests/test_gpu_plugin_api.py tests/test_gpu_planner.py tests/test_gpu_goal_scoring.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
( time timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py $TAG < gpurun_out/bench_n1_$TAG.json
tail -5 gpurun_out/bench_n1_$TAG.err
