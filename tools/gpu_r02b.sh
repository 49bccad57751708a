#!/bin/bash
# r02b: new GPU tests (reference classes, config shapes), bench line with config blocks + plugin e2e
TAG=r02b
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=12 -x 2>&1 | tail -60 > gpurun_out/pytest_gpu_$TAG.log
tail -8 gpurun_out/pytest_gpu_$TAG.log
( time python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py $TAG < gpurun_out/bench_n1_$TAG.json
tail -5 gpurun_out/bench_n1_$TAG.err
