#!/bin/bash
# r02v (2 GPUs): plugin-call profile on rank-less single process, then the N=2 bench line (NUMA binding, NCCL log)
TAG=r02v
mkdir -p gpurun_out
python tools/plugin_probe.py > gpurun_out/plugin_probe_$TAG.txt 2>&1; head -75 gpurun_out/plugin_probe_$TAG.txt
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1; head -12 gpurun_out/topo_$TAG.txt
lscpu | grep -E "NUMA|Socket|^CPU\(s\)" 
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_$TAG.json 2> gpurun_out/bench_n2_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py n2 < gpurun_out/bench_n2_$TAG.json
grep -E "numa binding|NCCL:|rank [0-9]: device" gpurun_out/bench_n2_$TAG.err | head -20
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_r02v.json').read().strip().splitlines()[-1])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_by_transfer_mode"], "plugin batch", d["e2e_plugin"]["batch"]["ms_per_call_max_over_ranks"], "allgather", d["allgather_ms"])
for k,v in d["configs"].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","allgather_ms","parity_frac_within_1e-4","error","batch_per_gpu","scenes_per_gpu","block_wall_s")})
print({k:v for k,v in d.get("nccl",{}).items() if k != "lines"}, d.get("numa_binding_rank0"))
PY
