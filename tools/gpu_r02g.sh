#!/bin/bash
TAG=r02g
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_goal_set.py -m gpu -q 2>&1 | tail -3
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
for C in 0 4 5 3; do
  OMGB_STEP_CONFIG=$C timeout 300 python bench.py $Q > gpurun_out/ab2_c2_cfg$C.json 2> gpurun_out/ab2_c2_cfg$C.err
  python tools/bench_summary.py c2_cfg$C < gpurun_out/ab2_c2_cfg$C.json
done
python tools/phase_profile.py > gpurun_out/phase_$TAG.txt 2>&1; tail -30 gpurun_out/phase_$TAG.txt
