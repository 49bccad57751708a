#!/bin/bash
# A/B bench of library variants built into omg_planner_b200/lib/var_*.so (diagnostic; run on the GPU box)
for rep in 1 2; do
for f in omg_planner_b200/lib/var_*.so; do
  OMGB_LIB=$PWD/$f python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>/dev/null | python tools/bench_summary.py "$(basename $f) $AB_TAG" | cut -c1-130
done
done
