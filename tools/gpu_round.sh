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
python tools/ncu_summary.py gpurun_out/chomp_full_$TAG.ncu-rep "chomp_step_kernel, bench.py --steps 3 --warmup 3 ($TAG)" > gpurun_out/ncu_chomp_$TAG.txt
for K in sdf_pack_entry point_sdf_kernel ik_chain_kernel; do
  ncu --set full --clock-control none -k regex:$K -s 1 -c 1 -f -o gpurun_out/aux_${K}_$TAG \
      python tools/bench_aux.py > gpurun_out/aux_ncu_${K}_$TAG.log 2>&1
  python tools/ncu_summary.py gpurun_out/aux_${K}_$TAG.ncu-rep "$K, tools/bench_aux.py ($TAG)" > gpurun_out/ncu_${K}_$TAG.txt
  rm -f gpurun_out/aux_${K}_$TAG.ncu-rep
done
python tools/bench_goalset_plan.py > gpurun_out/goalset_plan_$TAG.json 2> gpurun_out/goalset_plan_$TAG.err
for K in learner_update_kernel goal_cost_kernel; do
  B=1024 ncu --set full --clock-control none -k regex:$K -s 10 -c 1 -f -o gpurun_out/gs_${K}_$TAG \
      python tools/bench_goalset_plan.py > gpurun_out/gs_ncu_${K}_$TAG.log 2>&1
  python tools/ncu_summary.py gpurun_out/gs_${K}_$TAG.ncu-rep "$K, B=1024 G=20 tools/bench_goalset_plan.py ($TAG)" > gpurun_out/ncu_${K}_$TAG.txt
  rm -f gpurun_out/gs_${K}_$TAG.ncu-rep
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
du -sh gpurun_out
ls -la gpurun_out | tail -12
