#!/bin/bash
# r02ak: two-steps-per-round bisection in the learner kernel for small batches (bit-identical; A/B by
# OMGB_LEARNER_TWO_STEP=0/1): full GPU suite, goal-set plan timings at 1 / 256 / 1024 trajectories, config-3 block A/B,
# launch list of a one-trajectory plan, the bench line of this binary, smoke
TAG=r02ak
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
for two in 0 1; do
OMGB_LEARNER_TWO_STEP=$two SKIP_HOST=1 REPS=3 timeout 300 python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_two${two}_$TAG.json 2> gpurun_out/goalset_plan_$TAG.err; echo two=$two; cut -c150-900 gpurun_out/goalset_plan_two${two}_$TAG.json
OMGB_LEARNER_TWO_STEP=$two B=256 SKIP_HOST=1 SKIP_SINGLE=1 REPS=3 timeout 300 python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_b256_two${two}_$TAG.json 2>> gpurun_out/goalset_plan_$TAG.err; cut -c150-400 gpurun_out/goalset_plan_b256_two${two}_$TAG.json
done
B=1 SKIP_HOST=1 SKIP_SINGLE=1 REPS=2 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_plan_b1_$TAG.csv python tools/bench_goalset_plan.py > gpurun_out/plan_b1_ncu_$TAG.log 2>&1
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open('gpurun_out/launches_plan_b1_r02ak.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[re.sub(r"\(.*", "", r[ki])].append(float(r[vi].replace(",", "")))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:5]:
    print("%-60s n=%4d  mean %.2f us  min %.2f max %.2f" % (k[:60], len(v), sum(v)/len(v)/1e3, min(v)/1e3, max(v)/1e3))
PY
OMGB_LEARNER_TWO_STEP=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-aux --no-plugin --configs=3 > gpurun_out/bench_c3_two0_$TAG.json 2> gpurun_out/bench_c3_two0_$TAG.err
( time timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py full_$TAG < gpurun_out/bench_n1_$TAG.json | cut -c1-200
python - <<'PY'
import json
for f in ('bench_c3_two0_r02ak', 'bench_n1_r02ak'):
    d=json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
    for k,v in d["configs"].items():
        print(f, k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","value_one_scene_at_a_time","ms_per_step_one_scene_at_a_time","parity_frac_within_1e-4","error","block_wall_s")})
PY
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
