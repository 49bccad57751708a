#!/bin/bash
# Regenerates every fixture under tests/golden/ from the reference's own code (BUILD CONTAINER ONLY: needs
# /root/reference).  Each script documents what it runs unmodified and what it stubs.
set -e
cd "$(dirname "$0")/.."
make -s -C oracle            # CPU restatements
make -s -C oracle ref        # the reference's KDL (oracle/_ref/libkdl_ik.so) and CUDA operator source (libsdf_ref.so)
python tools/make_golden_sdf_interp.py # sdf_interp.npz : the reference's getValueInterpolated / getGradientInterpolated
python tools/make_golden.py            # chomp_*.npz    : omg.cost.Cost + omg.optimizer.Optimizer, 7 modes
python tools/make_golden_learner.py    # learner_*.npz  : omg.online_learner.Learner in Planner.plan's interleave
python tools/make_golden_plan.py       # plan_*.npz     : omg.planner.Planner.plan
python tools/make_golden_assets.py     # assets_*.npz   : util.interpolate_waypoints, Trajectory, SignedDensityField, combine_sdfs, PointEnv
python tools/make_golden_goalset.py    # ik_kdl.npz, goalset_*.npz : KDL inverse kinematics, Planner goal-set construction
ls -la tests/golden
