#!/bin/bash
TAG=r02e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
OMGB_SDF_LAYOUT=1 timeout 900 python -m pytest tests/test_gpu_chomp_step.py tests/test_gpu_edge_cases.py tests/test_gpu_goal_scoring.py tests/test_gpu_planner.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu_quad_$TAG.log
tail -2 gpurun_out/pytest_gpu_quad_$TAG.log
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
for V in "A=1" "OMGB_NO_BULK=1" "OMGB_SDF_LAYOUT=1"; do
  env $V timeout 300 python bench.py $Q > gpurun_out/ab_c2_$V.json 2> gpurun_out/ab_c2_$V.err
  python tools/bench_summary.py c2_$V < gpurun_out/ab_c2_$V.json
  env $V timeout 300 python bench.py $Q --waypoints 60 --objects 20 --grid 256 --steps 10 > gpurun_out/ab_c4_$V.json 2> gpurun_out/ab_c4_$V.err
  python tools/bench_summary.py c4_$V < gpurun_out/ab_c4_$V.json
done
