#!/bin/bash
# One GPU-box session: tests, bench lines, launch list, ncu captures.  Outputs under gpurun_out/.
TAG=${1:-r01e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 > gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
python tools/bench_summary.py $TAG < gpurun_out/bench_n1_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/b_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chomp_step -s 6 -c 1 -f -o gpurun_out/chomp_full_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/b_ncu_$TAG.log 2>&1
ncu --set full --clock-control none -k regex:"sdf_pack|ik_chain|point_sdf" -c 6 -f -o gpurun_out/aux_full_$TAG \
    python tools/bench_aux.py > gpurun_out/aux_ncu_$TAG.log 2>&1
ls -la gpurun_out | tail -12
