#!/usr/bin/env python
"""Goal-set plans with the online learner (the reference's default mode: goal_set_proj, ol_alg MD), whole Planner.plan
wall time for a batch: learner update on the host (north_star's split) vs device-resident pipeline
(omgb_goal_costs -> omgb_learner_update -> omgb_chomp_plan_step).  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch

    import helpers as H
    from omg_planner_b200 import core as C
    from omg_planner_b200 import scene as S
    from omg_planner_b200.config import ChompConfig
    from omg_planner_b200.planner import Planner
    from omg_planner_b200.robot import PandaConstants

    B, G = int(os.environ.get("B", 1024)), int(os.environ.get("G", 20))
    sc = S.make_scene(num_objects=10, grid=128, seed=0)
    robot = PandaConstants()
    goals, reach = S.make_goal_sets(B, G, robot.joint_lower_limit, robot.joint_upper_limit, seed=4, spread=0.3)
    res = {"workload": "%d trajectories x %d goals x 30 waypoints, 10 SDFs @128^3, ol_alg MD, standoff, 50 + 20 iterations, "
                       "pre_terminate off (every iteration runs)" % (B, G)}
    reps = int(os.environ.get("REPS", 2))
    prof = os.environ.get("PLAN_PROFILE")          # cProfile of the last device-learner plan -> this file
    for host in ((False,) if os.environ.get("SKIP_HOST") else (True, False)):
        cfg = ChompConfig(goal_set_proj=True, use_standoff=True, ol_alg="MD", pre_terminate=False, host_learner=host)
        env = H.make_env(sc, cfg, robot)
        target = env.objects[env.target_idx]
        target.grasps, target.reach_grasps = goals, reach
        traj = C.Trajectory(30, cfg=cfg, start=np.tile(S.START_CONF, (B, 1)), end=goals[:, 0])
        planner = Planner(env, traj)
        for rep in range(reps):   # later passes = warm
            traj = C.Trajectory(30, cfg=cfg, start=np.tile(S.START_CONF, (B, 1)), end=goals[:, 0])
            traj.goal_set = goals
            planner.update(env, traj)
            torch.cuda.synchronize()
            pr = None
            if prof and not host and rep == reps - 1:
                import cProfile
                pr = cProfile.Profile()
                pr.enable()
            t0 = time.perf_counter()
            planner.plan(traj)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if pr is not None:
                import io
                import pstats
                pr.disable()
                buf = io.StringIO()
                pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(35)
                open(prof, "w").write("plan wall %.4f s\n" % dt + buf.getvalue())
        iters = cfg.optim_steps + cfg.extra_smooth_steps
        res["host_learner" if host else "device_learner"] = {
            "plan_wall_s": dt, "ms_per_iteration": dt / iters * 1e3, "trajectory_iterations_per_s": B * iters / dt,
            "goals_selected": int(len(set(np.array(planner.selected_goals).reshape(-1).tolist())))}
    if "host_learner" in res:
        res["speedup"] = res["host_learner"]["plan_wall_s"] / res["device_learner"]["plan_wall_s"]
    if os.environ.get("SKIP_SINGLE"):
        print(json.dumps(res))
        return
    # one trajectory (the reference's own shape): whole Planner.plan latency, fixed goal (one persistent launch) and
    # goal set with the MD learner (device pipeline)
    one = {}
    for name, kw in (("fixed_goal", dict(goal_set_proj=False)), ("goal_set_md", dict(goal_set_proj=True, ol_alg="MD"))):
        cfg = ChompConfig(use_standoff=True, pre_terminate=False, **kw)
        env = H.make_env(sc, cfg, robot)
        target = env.objects[env.target_idx]
        target.grasps, target.reach_grasps = goals[0], reach[0]
        traj = C.Trajectory(30, cfg=cfg, start=S.START_CONF, end=goals[0, 0])
        planner = Planner(env, traj)
        ts = []
        sprof = os.environ.get("SINGLE_PROFILE")       # cProfile of one more single-trajectory goal-set plan -> this file
        for rep in range(5 if (sprof and name == "goal_set_md") else 4):
            traj = C.Trajectory(30, cfg=cfg, start=S.START_CONF, end=goals[0, 0])
            traj.goal_set = goals[0]
            planner.update(env, traj)
            torch.cuda.synchronize()
            pr = None
            if rep == 4:
                import cProfile
                pr = cProfile.Profile()
                pr.enable()
            t0 = time.perf_counter()
            planner.plan(traj)
            torch.cuda.synchronize()
            dt1 = time.perf_counter() - t0
            if pr is not None:
                import io
                import pstats
                pr.disable()
                buf = io.StringIO()
                pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(30)
                open(sprof, "w").write("single-trajectory goal-set plan, wall %.4f s (under cProfile)\n" % dt1 + buf.getvalue())
            else:
                ts.append(dt1)
        one[name] = {"plan_wall_ms": min(ts[1:]) * 1e3, "iterations": cfg.optim_steps + cfg.extra_smooth_steps}
        # the plugin call the reference's own loop makes: Optimizer.optimize(traj, force_update=True), one trajectory
        t0 = time.perf_counter()
        for _ in range(50):
            planner.optim.optimize(traj, force_update=True)
        torch.cuda.synchronize()
        one[name]["optimize_call_ms"] = (time.perf_counter() - t0) / 50 * 1e3
    res["single_trajectory_plan"] = one
    print(json.dumps(res))


if __name__ == "__main__":
    main()
