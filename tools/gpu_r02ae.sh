#!/bin/bash
# r02ae: validation of the round's final state on one B200: full GPU suite, reference arm, full bench line (with the
# config blocks), launch list of the bench command, ncu --set full of the step kernel (config 2, config-4 shape) and of
# the goal kernel, source-level hot lines, smoke
TAG=r02ae
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 300 gpurun_out/bench_ref_$TAG.json
( time timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py $TAG < gpurun_out/bench_n1_$TAG.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r02ae.json').read().strip().splitlines()[-1])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_by_transfer_mode"])
print("plugin", d["e2e_plugin"]["single"]["ms_per_call"], d["e2e_plugin"]["batch"]["ms_per_call"])
for k,v in d["configs"].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","parity_frac_within_1e-4","error","value_one_scene_at_a_time","block_wall_s")})
    if "plan_persistent" in v: print("   plan", v["plan_persistent"])
print({k:(v.get("ms"), v.get("chains_per_s")) for k,v in d["aux_kernels"].items()})
PY
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 $Q > gpurun_out/b_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chomp_step -s 8 -c 1 -f -o gpurun_out/chomp_full_c2_$TAG \
    python bench.py --steps 3 --warmup 3 $Q > gpurun_out/b_ncu_c2_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/chomp_full_c2_$TAG.ncu-rep "chomp_step_kernel, bench.py --steps 3 --warmup 3 (config 2: 1024 x 30 wpt, 10 SDFs @128^3) ($TAG)" > gpurun_out/ncu_chomp_c2_$TAG.txt
ncu -i gpurun_out/chomp_full_c2_$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/chomp_c2_src_$TAG.csv 2>/dev/null
python tools/ncu_lines.py gpurun_out/chomp_c2_src_$TAG.csv 60 > gpurun_out/ncu_chomp_c2_lines_$TAG.txt 2>&1
rm -f gpurun_out/chomp_c2_src_$TAG.csv
ncu --set full --clock-control none -k regex:chomp_step -s 8 -c 1 -f -o gpurun_out/chomp_full_c4_$TAG \
    python bench.py --steps 3 --warmup 3 $Q --waypoints 60 --objects 20 --grid 256 > gpurun_out/b_ncu_c4_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/chomp_full_c4_$TAG.ncu-rep "chomp_step_kernel, bench.py --steps 3 --warmup 3 --waypoints 60 --objects 20 --grid 256 (config-4 shape per GPU at N=8: 1024 x 60 wpt, 20 SDFs @256^3) ($TAG)" > gpurun_out/ncu_chomp_c4_$TAG.txt
rm -f gpurun_out/chomp_full_c4_$TAG.ncu-rep
B=1024 SKIP_HOST=1 SKIP_SINGLE=1 ncu --set full --clock-control none -k regex:goal_cost_kernel -s 10 -c 1 -f -o gpurun_out/gs_goal_cost_$TAG \
    python tools/bench_goalset_plan.py > gpurun_out/gs_ncu_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/gs_goal_cost_$TAG.ncu-rep "goal_cost_kernel, B=1024 G=20 tools/bench_goalset_plan.py ($TAG)" > gpurun_out/ncu_goal_cost_$TAG.txt
rm -f gpurun_out/gs_goal_cost_$TAG.ncu-rep
REPS=4 SKIP_HOST=1 SKIP_SINGLE=1 python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_$TAG.json 2> gpurun_out/goalset_plan_$TAG.err; tail -c 600 gpurun_out/goalset_plan_$TAG.json
python tools/bench_goal_scoring.py > gpurun_out/goal_scoring_$TAG.json 2>/dev/null; tail -c 500 gpurun_out/goal_scoring_$TAG.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
head -8 gpurun_out/ncu_chomp_c2_$TAG.txt; grep -E "duration|warps_active|issue_active|dram__bytes" gpurun_out/ncu_chomp_c4_$TAG.txt gpurun_out/ncu_goal_cost_$TAG.txt
du -sh gpurun_out
