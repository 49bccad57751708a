#!/bin/bash
# r02ag: staging mbarrier initialised once per CTA (phase parity per item, proxy fence before a persistent CTA's next bulk
# copies): racecheck over every entry point and the heavy-trajectory target, full GPU suite, short bench line, smoke
TAG=r02ag
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/racecheck_smoke_$TAG.log 2>&1; grep -v "^{'" gpurun_out/racecheck_smoke_$TAG.log | tail -6
OMGB_STEP_CONFIG=0 B=192 ITERS=2 timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_step_heavy.py > gpurun_out/racecheck_heavy_$TAG.log 2>&1; tail -3 gpurun_out/racecheck_heavy_$TAG.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/memcheck_smoke_$TAG.log 2>&1; tail -2 gpurun_out/memcheck_smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
for rep in 1 2; do
python bench.py --steps 20 --warmup 8 $Q > gpurun_out/bench_short_${rep}_$TAG.json 2> gpurun_out/bench_short_$TAG.err
python tools/bench_summary.py $TAG < gpurun_out/bench_short_${rep}_$TAG.json | cut -c1-160
done
python bench.py --steps 10 --warmup 8 $Q --waypoints 60 --objects 20 --grid 256 > gpurun_out/bench_short_c4_$TAG.json 2>> gpurun_out/bench_short_$TAG.err
python tools/bench_summary.py c4_$TAG < gpurun_out/bench_short_c4_$TAG.json | cut -c1-160
( time timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py full_$TAG < gpurun_out/bench_n1_$TAG.json | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
