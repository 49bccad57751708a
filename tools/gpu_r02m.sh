#!/bin/bash
# r02m: full GPU test suite, full bench line (N=1), launch list, ncu full captures (config 2, config-4 shape at
# 1024 and 8192 trajectories, goal cost kernel), goal-set plan bench
TAG=r02m
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
( time timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py $TAG < gpurun_out/bench_n1_$TAG.json
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 $Q > gpurun_out/b_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chomp_step -s 8 -c 1 -f -o gpurun_out/chomp_full_c2_$TAG \
    python bench.py --steps 3 --warmup 3 $Q > gpurun_out/b_ncu_c2_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/chomp_full_c2_$TAG.ncu-rep "chomp_step_kernel, bench.py --steps 3 --warmup 3 (config 2: 1024 x 30 wpt, 10 SDFs @128^3) ($TAG)" > gpurun_out/ncu_chomp_c2_$TAG.txt
ncu --set full --clock-control none --import-source on -k regex:chomp_step -s 8 -c 1 -f -o gpurun_out/chomp_full_c4_$TAG \
    python bench.py --steps 3 --warmup 3 $Q --waypoints 60 --objects 20 --grid 256 > gpurun_out/b_ncu_c4_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/chomp_full_c4_$TAG.ncu-rep "chomp_step_kernel, bench.py --steps 3 --warmup 3 --waypoints 60 --objects 20 --grid 256 (config-4 shape per GPU at N=8: 1024 x 60 wpt, 20 SDFs @256^3) ($TAG)" > gpurun_out/ncu_chomp_c4_$TAG.txt
ncu --set full --clock-control none -k regex:chomp_step -s 8 -c 1 -f -o gpurun_out/chomp_full_c4b_$TAG \
    python bench.py --steps 3 --warmup 3 $Q --batch 8192 --waypoints 60 --objects 20 --grid 256 > gpurun_out/b_ncu_c4b_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/chomp_full_c4b_$TAG.ncu-rep "chomp_step_kernel, bench.py --batch 8192 --waypoints 60 --objects 20 --grid 256 (config 4 on one GPU: 8192 x 60 wpt) ($TAG)" > gpurun_out/ncu_chomp_c4b_$TAG.txt
rm -f gpurun_out/chomp_full_c4b_$TAG.ncu-rep
python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_$TAG.json 2> gpurun_out/goalset_plan_$TAG.err
B=1024 ncu --set full --clock-control none -k regex:goal_cost_kernel -s 10 -c 1 -f -o gpurun_out/gs_goal_cost_$TAG \
    python tools/bench_goalset_plan.py > gpurun_out/gs_ncu_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/gs_goal_cost_$TAG.ncu-rep "goal_cost_kernel, B=1024 G=20 tools/bench_goalset_plan.py ($TAG)" > gpurun_out/ncu_goal_cost_$TAG.txt
rm -f gpurun_out/gs_goal_cost_$TAG.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
head -8 gpurun_out/ncu_chomp_c2_$TAG.txt; grep -E "duration|warps_active|issue_active|dram__bytes" gpurun_out/ncu_chomp_c4_$TAG.txt gpurun_out/ncu_goal_cost_$TAG.txt
du -sh gpurun_out
