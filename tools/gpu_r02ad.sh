#!/bin/bash
# r02ad: variant a = sphere cull + segmented winners + member cost from cached potentials + one-barrier block sums, plain
# histogram atomics; variant b (the built library) = a + per-warp member lists for the cost of the members.
# Full GPU suite on b, A/B against the round-start build at config 2 and the config-4 shape, phase profile of b.
TAG=r02ad
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
B=$PWD/omg_planner_b200/lib/libomgb200.so
A=$PWD/omg_planner_b200/lib/var_a.so
OLD=$PWD/omg_planner_b200/lib/var_old.so
for rep in 1 2; do
  for V in old a b; do
    L=$B; [ $V = old ] && L=$OLD; [ $V = a ] && L=$A
    OMGB_LIB=$L timeout 300 python bench.py --steps 20 --warmup 8 $Q > gpurun_out/ab_c2_${V}_$TAG.json 2> gpurun_out/ab_c2_${V}_$TAG.err
    python tools/bench_summary.py c2_$V < gpurun_out/ab_c2_${V}_$TAG.json | cut -c1-150
  done
done
for V in old a b; do
  L=$B; [ $V = old ] && L=$OLD; [ $V = a ] && L=$A
  OMGB_LIB=$L timeout 300 python bench.py $Q --waypoints 60 --objects 20 --grid 256 --steps 10 --warmup 8 > gpurun_out/ab_c4_${V}_$TAG.json 2> gpurun_out/ab_c4_${V}_$TAG.err
  python tools/bench_summary.py c4_$V < gpurun_out/ab_c4_${V}_$TAG.json | cut -c1-150
done
rm -f gpurun_out/phase_profile.txt
python tools/phase_profile.py > gpurun_out/phase_c2_$TAG.txt 2>&1; tail -24 gpurun_out/phase_c2_$TAG.txt | cut -c1-200
