#!/bin/bash
# r02ah (2 GPUs): the N=2 bench line of the final binary (config 2 per rank; configs 4 / 5 / 3 split over the ranks)
export TAG=${TAG:-r02ah}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_$TAG.json 2> gpurun_out/bench_n2_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py n2 < gpurun_out/bench_n2_$TAG.json | cut -c1-200
python - <<'PY'
import json, os
d=json.loads(open('gpurun_out/bench_n2_%s.json' % os.environ['TAG']).read().strip().splitlines()[-1])
print("e2e", d["e2e"]["ms_per_step"], "plugin batch", d["e2e_plugin"]["batch"]["ms_per_call_max_over_ranks"], "allgather", d["allgather_ms"])
for k,v in d["configs"].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","allgather_ms","parity_frac_within_1e-4","error","batch_per_gpu","scenes_per_gpu","block_wall_s")})
    if "plan_persistent" in v: print("   plan", v["plan_persistent"])
print({k:v for k,v in d.get("nccl",{}).items() if k != "lines"})
PY
