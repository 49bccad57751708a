#!/bin/bash
TAG=r02l
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_$TAG.json 2> gpurun_out/bench_n2_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py n2 < gpurun_out/bench_n2_$TAG.json
grep -E "c4dbg|rank" gpurun_out/bench_n2_$TAG.err | head
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_r02l.json').read().strip().splitlines()[-1])
for k,v in d["configs"].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","allgather_ms","parity_frac_within_1e-4","error","batch_per_gpu","scenes_per_gpu","block_wall_s")})
print(d.get("nccl"))
PY
ls -la gpurun_out | grep nccl | head -4
