#!/bin/bash
# r02al: parity at scale with the round's FINAL binary: device trajectories of 70-iteration plans at the
# config-2 / 4 / 5 shapes and whole goal-set plans with goal switching, dumped here, compared with the oracle in the
# build container (tools/parity_report.py check, tools/parity_report_goalset.py check)
TAG=r02al
mkdir -p gpurun_out
PARITY_SHAPE=config2 PARITY_MODES=fixed_topk,fixed_full,goalset_standoff_topk,goalset_single_full python tools/parity_report.py dump gpurun_out/pr_config2_$TAG.npz 2>&1 | tail -1
PARITY_SHAPE=config4 python tools/parity_report.py dump gpurun_out/pr_config4_$TAG.npz 2>&1 | tail -1
PARITY_SHAPE=config5 python tools/parity_report.py dump gpurun_out/pr_config5_$TAG.npz 2>&1 | tail -1
python tools/parity_report_goalset.py dump gpurun_out/pr_goalset_$TAG.npz 2>&1 | tail -1
ls -la gpurun_out/*.npz
