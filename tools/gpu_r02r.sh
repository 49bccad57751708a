#!/bin/bash
# r02r: history as one D2H (no concatenate), Ti by bincount: tests, goal-set plan bench, cProfile of the plan, launch list
TAG=r02r
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_goal_scoring.py tests/test_gpu_learner_device.py tests/test_gpu_planner.py tests/test_gpu_configs_fullsize.py tests/test_gpu_reference_classes.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log
tail -8 gpurun_out/pytest_gpu_$TAG.log
REPS=4 python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_$TAG.json 2> gpurun_out/goalset_plan_$TAG.err; cat gpurun_out/goalset_plan_$TAG.json
SKIP_HOST=1 SKIP_SINGLE=1 REPS=4 PLAN_PROFILE=gpurun_out/plan_cprofile_b1024_$TAG.txt python tools/bench_goalset_plan.py > gpurun_out/gsp2_$TAG.json 2>&1; cat gpurun_out/gsp2_$TAG.json
head -60 gpurun_out/plan_cprofile_b1024_$TAG.txt
B=256 SKIP_HOST=1 SKIP_SINGLE=1 REPS=4 PLAN_PROFILE=gpurun_out/plan_cprofile_b256_$TAG.txt python tools/bench_goalset_plan.py > gpurun_out/gsp3_$TAG.json 2>&1; cat gpurun_out/gsp3_$TAG.json
head -45 gpurun_out/plan_cprofile_b256_$TAG.txt
SKIP_HOST=1 SKIP_SINGLE=1 REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_plan_$TAG.csv \
    python tools/bench_goalset_plan.py > gpurun_out/gsp4_$TAG.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_plan_r02r.csv')) if len(r) > 10 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Metric Unit, Metric Value
half = len(rows) // 2
agg = collections.OrderedDict()
for r in rows[half:]:
    name = r[4].split('(')[0][:60]
    v = float(r[-1].replace(',', ''))
    unit = r[-2]
    v = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3 if unit in ('ms', 'msecond') else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
print("second plan: per-kernel launches and summed duration (us), serialised under ncu")
for k, (n, t) in agg.items(): print("%-62s %5d %10.1f" % (k, n, t))
print("total us", sum(t for n, t in agg.values()))
PY
