#!/bin/bash
# N=2: the distributed paths of bench.py (config blocks sharded over ranks, timed all-gather, NCCL log)
TAG=r02i
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_$TAG.json 2> gpurun_out/bench_n2_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py n2 < gpurun_out/bench_n2_$TAG.json
grep -E "NCCL:|Error|error|Traceback" gpurun_out/bench_n2_$TAG.err | head -10
ls gpurun_out/nccl_* | head
