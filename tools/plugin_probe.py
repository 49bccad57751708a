#!/usr/bin/env python
"""cProfile of the plugin call Optimizer.optimize(traj, force_update=True) over Cost.evaluate (diagnostic; run on the GPU
box): where the host time of one call goes, one trajectory and a batch of 1024."""
import cProfile
import io
import os
import pstats
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from omg_planner_b200 import scene as S
    from omg_planner_b200.robot import PandaConstants

    a = types.SimpleNamespace(waypoints=30, steps=50)
    mode = bench.DEFAULT_MODE
    scene = S.make_scene(num_objects=10, grid=128, seed=0)
    robot = PandaConstants()
    xi0, st, en, tails = S.make_trajectories(1024, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=0)
    prs = {2: cProfile.Profile(), 3: cProfile.Profile()}   # by traj.data.ndim: one trajectory / a batch
    from omg_planner_b200 import optimizer as O

    orig = O.Optimizer.optimize
    seen = {2: 0, 3: 0}

    def wrapped(self, traj, **kw):
        k = np.ndim(traj.data)
        seen[k] += 1
        if seen[k] > 3:          # (the warm-up calls stay out)
            prs[k].enable()
        try:
            return orig(self, traj, **kw)
        finally:
            prs[k].disable()

    O.Optimizer.optimize = wrapped
    res = bench.run_plugin_e2e(a, scene, mode, xi0, st, en, tails, robot, steps=50)
    O.Optimizer.optimize = orig
    for k, name in ((3, "batch"), (2, "single")):
        buf = io.StringIO()
        pstats.Stats(prs[k], stream=buf).sort_stats("cumulative").print_stats(30)
        print("=====", name, res[name])
        print(buf.getvalue()[:5000])


if __name__ == "__main__":
    main()
