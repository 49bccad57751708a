#!/bin/bash
# r02a: GPU tests incl. the reference-source operator pin; config-4 true-shape baseline (before the kernel work)
TAG=r02a
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 > gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
python bench.py --no-cpu-baseline --no-aux > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
python tools/bench_summary.py $TAG < gpurun_out/bench_n1_$TAG.json
timeout 900 python bench.py --batch 1024 --waypoints 60 --objects 20 --grid 256 --no-cpu-baseline --no-aux --steps 10 \
    > gpurun_out/bench_config4_$TAG.json 2> gpurun_out/bench_config4_$TAG.err
python tools/bench_summary.py ${TAG}_c4 < gpurun_out/bench_config4_$TAG.json
