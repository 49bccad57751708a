#!/bin/bash
# r02ai: omgb_chomp_plan_goalset (the goal-set plan's loop inside the library, one call per plan): full GPU suite,
# goal-set plan timings (1024 / 256 / 1 trajectories), the bench line of this binary, smoke
TAG=r02ai
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
SKIP_HOST=1 REPS=3 SINGLE_PROFILE=gpurun_out/single_plan_cprofile_$TAG.txt timeout 300 python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_$TAG.json 2> gpurun_out/goalset_plan_$TAG.err; cut -c1-900 gpurun_out/goalset_plan_$TAG.json
B=256 SKIP_HOST=1 SKIP_SINGLE=1 REPS=3 timeout 300 python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_b256_$TAG.json 2>> gpurun_out/goalset_plan_$TAG.err; cut -c1-400 gpurun_out/goalset_plan_b256_$TAG.json
head -25 gpurun_out/single_plan_cprofile_$TAG.txt | cut -c1-150
( time timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py full_$TAG < gpurun_out/bench_n1_$TAG.json | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r02ai.json').read().strip().splitlines()[-1])
for k,v in d["configs"].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","value_one_scene_at_a_time","ms_per_step_one_scene_at_a_time","parity_frac_within_1e-4","error","block_wall_s")})
PY
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
