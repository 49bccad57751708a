#!/bin/bash
# r02aj: where a ONE-trajectory goal-set plan's 95 us per iteration go: ncu launch list (kernel durations, serialised)
TAG=r02aj
mkdir -p gpurun_out
B=1 SKIP_HOST=1 SKIP_SINGLE=1 REPS=2 timeout 400 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -c 700 --csv --log-file gpurun_out/launches_plan_b1_$TAG.csv python tools/bench_goalset_plan.py > gpurun_out/plan_b1_ncu_$TAG.log 2>&1
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open('gpurun_out/launches_plan_b1_r02aj.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ii], {"k": r[ki]})[r[mi]] = r[vi]
agg = collections.defaultdict(list)
for v in per.values():
    name = re.sub(r"\(.*", "", v["k"])
    agg[name].append(float(v.get("gpu__time_duration.sum", "0").replace(",", "")))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v2 = v[len(v)//2:]   # second plan (warm)
    print("%-60s n=%4d  total %.1f us  mean %.2f us  min %.2f max %.2f (second half mean %.2f)" % (k[:60], len(v), sum(v)/1e3, sum(v)/len(v)/1e3, min(v)/1e3, max(v)/1e3, sum(v2)/len(v2)/1e3))
PY
