#!/bin/bash
# r02z: N ranks, the driver's scaling run at the round's final state (+ topology of the box)
N=${1:-8}
TAG=r02z
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n${N}_$TAG.txt 2>&1; lscpu | grep -E "NUMA|Socket|^CPU\(s\)|Model name" >> gpurun_out/topo_n${N}_$TAG.txt; head -14 gpurun_out/topo_n${N}_$TAG.txt | cut -c1-160
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py n$N < gpurun_out/bench_n${N}_$TAG.json
grep -E "numa binding|rank [0-9]: device|NCCL:.*nranks" gpurun_out/bench_n${N}_$TAG.err | head -24
python - $N <<'PY'
import json, sys
N=sys.argv[1]
d=json.loads(open('gpurun_out/bench_n%s_r02z.json' % N).read().strip().splitlines()[-1])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_by_transfer_mode"], "plugin batch", d["e2e_plugin"]["batch"]["ms_per_call_max_over_ranks"], "allgather", d["allgather_ms"])
for k,v in d["configs"].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","allgather_ms","parity_frac_within_1e-4","error","batch_per_gpu","scenes_per_gpu","block_wall_s","value_one_scene_at_a_time")})
    if "plan_persistent" in v: print("   plan_persistent", v["plan_persistent"])
print({k:v for k,v in d.get("nccl",{}).items() if k != "lines"}, d.get("numa_binding_rank0"))
PY
