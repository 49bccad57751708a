#!/bin/bash
# r02s: several goals per CTA for short lines; zero-copy bandwidth probe
TAG=r02s
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_goal_scoring.py tests/test_gpu_learner_device.py tests/test_gpu_planner.py tests/test_gpu_configs_fullsize.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log
tail -8 gpurun_out/pytest_gpu_$TAG.log
REPS=4 SKIP_HOST=1 python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_$TAG.json 2> gpurun_out/goalset_plan_$TAG.err; cat gpurun_out/goalset_plan_$TAG.json
REPS=4 SKIP_HOST=1 SKIP_SINGLE=1 OMGB_GOAL_GPC=1 python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_gpc1_$TAG.json 2> gpurun_out/goalset_plan_gpc1_$TAG.err; cat gpurun_out/goalset_plan_gpc1_$TAG.json
B=256 REPS=4 SKIP_HOST=1 SKIP_SINGLE=1 python tools/bench_goalset_plan.py 2>&1 | tail -1
python tools/bench_goal_scoring.py 2>/dev/null | tail -1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/zc tools/zerocopy_probe.cu && /tmp/zc > gpurun_out/zerocopy_probe_$TAG.txt 2>&1; cat gpurun_out/zerocopy_probe_$TAG.txt
SKIP_HOST=1 SKIP_SINGLE=1 REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_plan_$TAG.csv \
    python tools/bench_goalset_plan.py > gpurun_out/gsp4_$TAG.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_plan_r02s.csv')) if len(r) > 10 and r[0].isdigit()]
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
