#!/bin/bash
TAG=r02j
mkdir -p gpurun_out
timeout 600 python tools/bench_configs.py config4 > gpurun_out/c4_n1_$TAG.json 2> gpurun_out/c4_n1_$TAG.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c4_n1_r02j.json').read().strip().splitlines()[-1])
for k,v in d.items(): print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","plan_persistent","error","traceback","batch_per_gpu")})
PY
OMGB_SDF_LAYOUT=0 timeout 600 python tools/bench_configs.py config4 > gpurun_out/c4_n1_plain_$TAG.json 2> gpurun_out/c4_n1_plain_$TAG.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c4_n1_plain_r02j.json').read().strip().splitlines()[-1])
for k,v in d.items(): print("plain", k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","plan_persistent","error","traceback","batch_per_gpu")})
PY
