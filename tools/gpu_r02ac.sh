#!/bin/bash
# r02ac: obstacle cost always summed in phase 4a from the cached potentials (none in the points phase), leader-bucket histogram aggregation; full GPU suite, A/B against the round-start build
# against the previous build (omg_planner_b200/lib/var_old.so) at config 2 and the config-4 shape, phase profiles
TAG=r02ac
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
NEW=$PWD/omg_planner_b200/lib/libomgb200.so
OLD=$PWD/omg_planner_b200/lib/var_old.so
for rep in 1 2; do
  for V in old new newnoseg; do
    L=$NEW; E="A=1"
    [ $V = old ] && L=$OLD
    [ $V = newnoseg ] && E="OMGB_WIN_SEG_MIN=100000"
    env $E OMGB_LIB=$L timeout 300 python bench.py --steps 20 --warmup 8 $Q > gpurun_out/ab_c2_${V}_$TAG.json 2> gpurun_out/ab_c2_${V}_$TAG.err
    python tools/bench_summary.py c2_$V < gpurun_out/ab_c2_${V}_$TAG.json | cut -c1-150
  done
done
for V in old new newnoseg; do
  L=$NEW; E="A=1"
  [ $V = old ] && L=$OLD
  [ $V = newnoseg ] && E="OMGB_WIN_SEG_MIN=100000"
  env $E OMGB_LIB=$L timeout 300 python bench.py $Q --waypoints 60 --objects 20 --grid 256 --steps 10 --warmup 8 > gpurun_out/ab_c4_${V}_$TAG.json 2> gpurun_out/ab_c4_${V}_$TAG.err
  python tools/bench_summary.py c4_$V < gpurun_out/ab_c4_${V}_$TAG.json | cut -c1-150
done
rm -f gpurun_out/phase_profile.txt
python tools/phase_profile.py > gpurun_out/phase_c2_$TAG.txt 2>&1; tail -24 gpurun_out/phase_c2_$TAG.txt | cut -c1-200
W=60 O=20 GRID=256 python tools/phase_profile.py > gpurun_out/phase_c4_$TAG.txt 2>&1; tail -24 gpurun_out/phase_c4_$TAG.txt | cut -c1-200
REPS=4 SKIP_HOST=1 SKIP_SINGLE=1 python tools/bench_goalset_plan.py 2>/dev/null | tail -1 | cut -c1-400
