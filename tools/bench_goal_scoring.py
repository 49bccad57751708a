#!/usr/bin/env python
"""Throughput of the fused goal-scoring kernel (omgb_goal_costs) on the config-2 stand-in scene, with
the oracle's restatement of Learner.cost_vector's device half timed beside it on one host core.  Diagnostic; run on
the GPU box; appends one JSON line to gpurun_out/goal_scoring.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omg_planner_b200 import scene as S  # noqa: E402
from omg_planner_b200.config import ChompConfig  # noqa: E402
from omg_planner_b200.engine import ChompEngine  # noqa: E402
from omg_planner_b200.robot import PandaConstants  # noqa: E402


def main():
    B = int(os.environ.get("GS_BATCH", 1024)); G = int(os.environ.get("GS_GOALS", 20)); n = 30
    sc = S.make_scene(num_objects=10, grid=128, seed=0)
    cfg = ChompConfig()
    robot = PandaConstants()
    eng = ChompEngine(robot=robot).load_scene(sc, cfg)
    xi, st, en, tails = S.make_trajectories(B, n, robot.joint_lower_limit, robot.joint_upper_limit, seed=0)
    goals, _ = S.make_goal_sets(B, G, robot.joint_lower_limit, robot.joint_upper_limit, seed=1)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x, g = dev(xi), dev(goals)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    res = {}
    for first in (0, 15, 24):
        for _ in range(3):
            out = eng.goal_costs(x, first, g, cfg.time_interval, 0)
        torch.cuda.synchronize()
        ms = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = eng.goal_costs(x, first, g, cfg.time_interval, 0); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms = float(np.median(ms))
        configs = B * G * (n - first)
        res["first_%d" % first] = {"ms": ms, "goal_lines_per_s": B * G / (ms * 1e-3), "configurations_per_s": configs / (ms * 1e-3)}
    # oracle on one core, a few trajectories
    from oracle import chomp_ref as R, learner_ref as LR
    rr, rcfg = R.PandaRef(), R.RefConfig()
    t0 = time.perf_counter(); nb = 2
    for b in range(nb):
        LR.collision_costs(rr, sc, rcfg, xi[b, 0], goals[b], n)
    cpu_s = (time.perf_counter() - t0) / nb
    line = {"workload": "goal scoring: %d trajectories x %d goals x <=%d waypoints, 10 SDFs @128^3" % (B, G, n),
            "gpu": res, "cpu_oracle_one_core": {"s_per_trajectory_first_0": cpu_s, "goal_lines_per_s": G / cpu_s}}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "goal_scoring.json"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
