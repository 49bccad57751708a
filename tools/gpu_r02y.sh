#!/bin/bash
# r02y: the fp64 no-FMA experiment: 70-iteration trajectories of the shipped build and of a build
# whose fp64 multiply-adds are all rounded twice (-DOMGB_NO_FP64_FMA -fmad=false), dumped for tools/parity_report.py check
TAG=r02y
mkdir -p gpurun_out
for SH in config4 config5 config2; do
  M=goalset_standoff_topk; [ $SH = config2 ] && M=fixed_full
  PARITY_SHAPE=$SH PARITY_MODES=$M python tools/parity_report.py dump gpurun_out/pr_${SH}_fma_$TAG.npz > /dev/null 2>&1
  OMGB_LIB=$PWD/omg_planner_b200/lib/libomgb200_nofma.so PARITY_SHAPE=$SH PARITY_MODES=$M python tools/parity_report.py dump gpurun_out/pr_${SH}_nofma_$TAG.npz > /dev/null 2>&1
done
ls -la gpurun_out/pr_*_$TAG.npz
