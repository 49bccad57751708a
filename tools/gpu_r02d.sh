#!/bin/bash
TAG=r02d
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
for C in 0 3 1; do
  OMGB_STEP_CONFIG=$C timeout 300 python bench.py $Q > gpurun_out/ab_c2_cfg$C.json 2> gpurun_out/ab_c2_cfg$C.err
  python tools/bench_summary.py c2_cfg$C < gpurun_out/ab_c2_cfg$C.json
done
for C in 2 1; do
  OMGB_STEP_CONFIG=$C timeout 300 python bench.py $Q --waypoints 60 --objects 20 --grid 128 --steps 10 > gpurun_out/ab_c4_cfg$C.json 2> gpurun_out/ab_c4_cfg$C.err
  python tools/bench_summary.py c4shape128_cfg$C < gpurun_out/ab_c4_cfg$C.json
done
