#!/bin/bash
# r02u: full GPU suite + full bench line (N=1) after the goal-kernel / plan-history / plugin pass-through changes
TAG=r02u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
( time timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py $TAG < gpurun_out/bench_n1_$TAG.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r02u.json').read().strip().splitlines()[-1])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_by_transfer_mode"])
print("plugin", d["e2e_plugin"]["single"]["ms_per_call"], d["e2e_plugin"]["batch"]["ms_per_call"])
for k,v in d["configs"].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","parity_frac_within_1e-4","error","value_one_scene_at_a_time","block_wall_s")})
PY
tail -3 gpurun_out/bench_n1_$TAG.err
