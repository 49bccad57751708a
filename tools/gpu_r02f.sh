#!/bin/bash
TAG=r02f
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_goal_set.py -m gpu -q -s 2>&1 | grep -E "status agreement|chains |final goals|passed|failed|Error|assert" | head -30 > gpurun_out/pytest_goalset_$TAG.log
cat gpurun_out/pytest_goalset_$TAG.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
