#!/bin/bash
# N ranks: the driver's scaling run
N=${1:-8}
TAG=r02n
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py n$N < gpurun_out/bench_n${N}_$TAG.json
grep -E "c4dbg|rank [0-9]: device|NCCL:" gpurun_out/bench_n${N}_$TAG.err | head -30
python - $N <<'PY'
import json, sys
N=sys.argv[1]
d=json.loads(open('gpurun_out/bench_n%s_r02n.json' % N).read().strip().splitlines()[-1])
print("e2e modes", d["e2e"]["ms_per_step_by_transfer_mode"], "plugin batch", d["e2e_plugin"]["batch"]["ms_per_call_max_over_ranks"], "allgather", d["allgather_ms"])
for k,v in d["configs"].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","allgather_ms","parity_frac_within_1e-4","error","batch_per_gpu","scenes_per_gpu","block_wall_s")})
print({k:v for k,v in d.get("nccl",{}).items() if k != "lines"})
PY
