#!/usr/bin/env python
"""Goal-set construction end to end (what Planner.__init__ does before the first CHOMP iteration,
omg/planner.py:103-114): grasp poses -> batched IK of every (pose, seed) chain -> wrist-flip augmentation ->
hand-rotation filter -> collision filter (fused batch_obstacle_cost) -> diversity filter -> sampling -> initial goal
and trajectory.  Wall time of the whole thing for 100 and 300 synthetic grasps; the reference prints this as
"IK init time" (seconds, 4-process PyKDL pool).  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def synthetic_grasps(rng, n, obj_pose, radius=0.12):
    out = []
    for _ in range(n):
        d = rng.normal(size=3); d[2] = abs(d[2]) + 0.3; d /= np.linalg.norm(d)
        z = -d
        x = np.cross(z, rng.normal(size=3)); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        T = np.eye(4)
        T[:3, 0], T[:3, 1], T[:3, 2] = x, y, z
        T[:3, 3] = obj_pose[:3, 3] + d * radius
        out.append(np.linalg.inv(obj_pose) @ T)
    return np.stack(out)


def main():
    import torch

    import helpers as H
    from omg_planner_b200 import core as C
    from omg_planner_b200 import scene as S
    from omg_planner_b200.config import ChompConfig
    from omg_planner_b200.planner import Planner
    from omg_planner_b200.robot import PandaConstants
    from oracle import kdl_ik_ref as K

    sc = S.make_scene(num_objects=10, grid=128, seed=0)
    robot = PandaConstants()
    res = {}
    for n_grasps in (100, 300):
        cfg = ChompConfig(goal_set_proj=True, use_standoff=True, ol_alg="MD", goal_idx=-1)
        ts = []
        for rep in range(3):
            env = H.make_env(sc, cfg, robot)
            for i, o in enumerate(env.objects):
                o.compute_grasp = i == env.target_idx
                o.grasp_potentials, o.grasp_vis_points, o.seeds = [], [], []
            target = env.objects[env.target_idx]
            target.grasps_poses = synthetic_grasps(np.random.RandomState(n_grasps), n_grasps, target.pose_mat)
            traj = C.Trajectory(30, cfg=cfg, start=S.START_CONF, end=S.START_CONF)
            np.random.seed(0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            planner = Planner(env, traj)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        entry = {"wall_ms": min(ts) * 1e3, "goals_kept": int(len(target.grasps)), "chains": (n_grasps - 1) * 13}
        # the reference's cost for the IK alone, extrapolated from a sample of chains on one core (its pool has 4)
        ch = K.PandaChain(robot.pose_0, robot.joint_lower_limit, robot.joint_upper_limit)
        if K.have_ref():
            from omg_planner_b200.ik import poses_to_targets
            from omg_planner_b200.goal_set import UTIL_ANCHOR_SEEDS
            poses = np.matmul(target.pose_mat, target.grasps_poses[:12])
            back = np.tile(np.eye(4), (6, 1, 1)); back[:, 2, 3] = -0.08 * np.array([4, 0, 1, 2, 3, 4]) / 5.0
            tg = poses_to_targets(np.matmul(poses[:, None], back[None]))
            seeds = np.concatenate([[S.START_CONF[:7]], UTIL_ANCHOR_SEEDS[:12, :7]])
            t0 = time.perf_counter()
            for p in range(12):
                for s in range(13):
                    q = seeds[s]
                    for t in range(6):
                        r, rc, raw = ch.ref_ik(tg[p, t, :3], tg[p, t, 3:], q)
                        if rc < 0:
                            break
                        q = r
            per_chain = (time.perf_counter() - t0) / (12 * 13)
            entry["reference_kdl_ik_only_ms_4_processes_extrapolated"] = per_chain * entry["chains"] / 4 * 1e3
        res["%d_grasps" % n_grasps] = entry
    print(json.dumps(res))


if __name__ == "__main__":
    main()
