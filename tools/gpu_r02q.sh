#!/bin/bash
# r02q: goal_cost_kernel with the warp-private exact-evaluation queue: parity tests, goal-set plan bench, ncu summary
TAG=r02q
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_goal_scoring.py tests/test_gpu_learner_device.py tests/test_gpu_planner.py tests/test_gpu_configs_fullsize.py tests/test_gpu_plugin_api.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log
tail -8 gpurun_out/pytest_gpu_$TAG.log
python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_$TAG.json 2> gpurun_out/goalset_plan_$TAG.err; cat gpurun_out/goalset_plan_$TAG.json
python tools/bench_goal_scoring.py > gpurun_out/goal_scoring_$TAG.json 2> gpurun_out/goal_scoring_$TAG.err; cat gpurun_out/goal_scoring_$TAG.json
B=1024 timeout 600 ncu --set full --clock-control none --import-source on -k regex:goal_cost_kernel -s 10 -c 1 -f -o gpurun_out/gs_goal_cost_$TAG \
    python tools/bench_goalset_plan.py > gpurun_out/gs_ncu_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/gs_goal_cost_$TAG.ncu-rep "goal_cost_kernel, B=1024 G=20 tools/bench_goalset_plan.py ($TAG)" > gpurun_out/ncu_goal_cost_$TAG.txt
cat gpurun_out/ncu_goal_cost_$TAG.txt
ncu -i gpurun_out/gs_goal_cost_$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/gs_goal_cost_src_$TAG.csv 2>/dev/null
python tools/ncu_lines.py gpurun_out/gs_goal_cost_src_$TAG.csv 40 > gpurun_out/ncu_goal_cost_lines_$TAG.txt
head -30 gpurun_out/ncu_goal_cost_lines_$TAG.txt
rm -f gpurun_out/gs_goal_cost_src_$TAG.csv
