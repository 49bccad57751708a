#!/bin/bash
# r02t: goal kernel with redux accumulation: goals-per-CTA A/B; config-3 block with host threads; config 5 block
TAG=r02t
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_goal_scoring.py tests/test_gpu_learner_device.py tests/test_gpu_planner.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
python tools/bench_goal_scoring.py 2>/dev/null | tail -1 | tee gpurun_out/goal_scoring_$TAG.json
OMGB_GOAL_GPC=1 python tools/bench_goal_scoring.py 2>/dev/null | tail -1 | tee gpurun_out/goal_scoring_gpc1_$TAG.json
REPS=5 SKIP_HOST=1 SKIP_SINGLE=1 python tools/bench_goalset_plan.py 2>/dev/null | tail -1 | tee gpurun_out/goalset_plan_$TAG.json
REPS=5 SKIP_HOST=1 SKIP_SINGLE=1 OMGB_GOAL_GPC=1 python tools/bench_goalset_plan.py 2>/dev/null | tail -1 | tee gpurun_out/goalset_plan_gpc1_$TAG.json
( time python tools/bench_configs.py config3 config5 > gpurun_out/configs35_$TAG.json 2> gpurun_out/configs35_$TAG.err ) 2>&1 | tail -3
python - <<'PY'
import json
d = json.loads(open('gpurun_out/configs35_r02t.json').read().strip().splitlines()[-1])
for k, v in d.items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ('workload', 'timing', 'parity', 'api')})
    print('   parity', v.get('parity'))
PY
tail -5 gpurun_out/configs35_$TAG.err
