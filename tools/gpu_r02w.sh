#!/bin/bash
# r02w: wide CTAs for the heaviest trajectories: bit-identity test, step parity tests, A/B bench (config 2), phase profile
TAG=r02w
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_chomp_step.py tests/test_gpu_edge_cases.py tests/test_gpu_planner.py tests/test_gpu_plugin_api.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
for W in -1 0; do
  OMGB_WIDE_CTAS=$W python bench.py --steps 20 --warmup 8 $Q > gpurun_out/ab_wide${W}_$TAG.json 2> gpurun_out/ab_wide${W}_$TAG.err
  python tools/bench_summary.py wide$W < gpurun_out/ab_wide${W}_$TAG.json
done
for W in -1 0; do
  OMGB_WIDE_CTAS=$W python bench.py --steps 20 --warmup 8 $Q > gpurun_out/ab2_wide${W}_$TAG.json 2> gpurun_out/ab2_wide${W}_$TAG.err
  python tools/bench_summary.py wide$W < gpurun_out/ab2_wide${W}_$TAG.json
done
rm -f gpurun_out/phase_profile.txt
OMGB_WIDE_CTAS=-1 python tools/phase_profile.py > gpurun_out/phase_c2_wide_$TAG.txt 2>&1; tail -22 gpurun_out/phase_c2_wide_$TAG.txt
REPS=4 SKIP_HOST=1 SKIP_SINGLE=1 python tools/bench_goalset_plan.py 2>/dev/null | tail -1
REPS=4 SKIP_HOST=1 SKIP_SINGLE=1 OMGB_WIDE_CTAS=0 python tools/bench_goalset_plan.py 2>/dev/null | tail -1
