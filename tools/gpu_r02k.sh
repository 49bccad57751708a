#!/bin/bash
TAG=r02k
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-aux --no-plugin --configs=4 --no-parity --steps 5"
timeout 600 python bench.py $Q > gpurun_out/k_n1.json 2> gpurun_out/k_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/k_n1.json').read().strip().splitlines()[-1])
print("N=1 headline", d["ms_per_step"], "config4", d["configs"]["config4"].get("ms_per_step"), d["configs"]["config4"].get("error"))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $Q > gpurun_out/k_n2.json 2> gpurun_out/k_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/k_n2.json').read().strip().splitlines()[-1])
print("N=2 headline", d["ms_per_step"], "config4", d["configs"]["config4"].get("ms_per_step"), d["configs"]["config4"].get("error"))
PY
grep -E "rank|c4dbg" gpurun_out/k_n2.err | head
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
