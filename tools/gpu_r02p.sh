#!/bin/bash
# r02p: sanity GPU tests after the container rebuild, e2e probe, goal_cost source-level profile, config-4-shape phases
TAG=r02p
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python tools/e2e_probe.py > gpurun_out/e2e_probe_$TAG.txt 2>&1; cat gpurun_out/e2e_probe_$TAG.txt | tail -12
rm -f gpurun_out/phase_profile.txt
W=60 O=20 GRID=256 timeout 300 python tools/phase_profile.py > gpurun_out/phase_c4_$TAG.txt 2>&1; tail -30 gpurun_out/phase_c4_$TAG.txt
B=1024 timeout 600 ncu --set full --clock-control none --import-source on -k regex:goal_cost_kernel -s 10 -c 1 -f -o gpurun_out/gs_goal_cost_$TAG \
    python tools/bench_goalset_plan.py > gpurun_out/gs_ncu_$TAG.log 2>&1
ncu -i gpurun_out/gs_goal_cost_$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/gs_goal_cost_src_$TAG.csv 2>/dev/null
python tools/ncu_lines.py gpurun_out/gs_goal_cost_src_$TAG.csv 60 > gpurun_out/ncu_goal_cost_lines_$TAG.txt
head -70 gpurun_out/ncu_goal_cost_lines_$TAG.txt
rm -f gpurun_out/gs_goal_cost_src_$TAG.csv
du -sh gpurun_out
