#!/usr/bin/env python
"""Generate tests/golden/ik_kdl.npz and tests/golden/goalset_*.npz.

  ik_kdl      single inverse-kinematics problems solved by the reference's OWN vendored KDL
              (oracle/_ref/libkdl_ik.so, compiled from /root/reference by oracle/kdl_ref/Makefile): targets, seeds,
              status codes and joint solutions.
  goalset_*   the UNMODIFIED reference Planner.solve_and_process_ik / setup_goal_set / grasp_init
              (omg/planner.py:187-597 under the stubs of tools/ref_harness.py) on synthetic grasp poses, with
              cfg.ROBOT.inverse_kinematics bound to that KDL library (the reference binds PyKDL, i.e. the same C++),
              the reference's own FK for the hand-rotation filter and its own Cost for the collision filter.

BUILD-CONTAINER ONLY (needs /root/reference).  The committed .npz files are what travels.
"""
import builtins
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as H  # noqa: E402
from omg_planner_b200 import scene as S  # noqa: E402
from oracle import chomp_ref as R  # noqa: E402
from oracle import kdl_ik_ref as K  # noqa: E402

SCENE_ARGS = dict(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
CASES = {
    "standoff_parallel": dict(use_standoff=True, ik_parallel=True, n_grasps=14, seed=1),
    "single_sequential": dict(use_standoff=False, ik_parallel=False, n_grasps=10, seed=2),
    # placement: the target is attached to the hand, one relative hand pose up-sampled by 50 rotations about the
    # object's z axis (omg/planner.py:324-335, 493-498); no wrist-flip augmentation / hand-rotation filter, reach tails
    # are not reversed, the table's collision parameters change (omg/cost.py:325-328)
    "placement_zupsample": dict(use_standoff=True, ik_parallel=False, n_grasps=1, seed=3, attached=True, z_upsample=True),
    # cfg.increment_iks (omg/config.py:94): extra IK seeds from earlier solutions -- ten random ones per pool group of
    # four poses (omg/planner.py:436-441), the closest earlier solution in the sequential loop (:365-373)
    "increment_parallel": dict(use_standoff=True, ik_parallel=True, n_grasps=11, seed=4, increment_iks=True),
    "increment_sequential": dict(use_standoff=False, ik_parallel=False, n_grasps=8, seed=5, increment_iks=True),
}


def synthetic_grasps(rng, n, obj_pose, radius=0.12):
    """Hand poses (object coordinates) looking at the object centre from the upper hemisphere: hand z = approach
    direction, random roll (stand-in for data/grasps/simulated/*.npy)."""
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
    assert K.build_ref() and K.have_ref(), "oracle/_ref/libkdl_ik.so could not be built"
    ns = H.load_reference()
    sys.modules.setdefault("ycb_render.ycb_renderer", H._Anything("ycb_render.ycb_renderer"))
    planner_mod = importlib.import_module("omg.planner")
    cfg = ns.cfg
    cfg.report_time = False
    cfg.silent = True
    robot = R.PandaRef()
    chain = K.PandaChain(robot.pose_0, robot.lower, robot.upper)
    out_dir = os.path.join(ROOT, "tests", "golden")

    # ---- single problems ------------------------------------------------------------------------------------
    rng = np.random.RandomState(21)
    P, Sd = 60, 4
    poses = np.stack([chain.ref_fk_hand(rng.uniform(chain.lo, chain.hi)) for _ in range(P)])
    poses[::6, :3, 3] += rng.uniform(-0.4, 0.4, (len(poses[::6]), 3))           # some unreachable
    from omg_planner_b200.ik import poses_to_targets
    targets = poses_to_targets(poses)
    seeds = np.concatenate([[S.START_CONF[:7]], rng.uniform(chain.lo, chain.hi, (Sd - 1, 7))])
    status = np.zeros((P, Sd), np.int32); sols = np.zeros((P, Sd, 7))
    for p in range(P):
        for s in range(Sd):
            _, status[p, s], sols[p, s] = chain.ref_ik(targets[p, :3], targets[p, 3:], seeds[s])
    np.savez_compressed(os.path.join(out_dir, "ik_kdl.npz"), targets=targets, seeds=seeds, status=status, sols=sols,
                        pose_0=robot.pose_0[:8], lower=chain.lo, upper=chain.hi)
    print("ik_kdl: %d problems, %d solved" % (P * Sd, (status >= 0).sum()))

    # ---- goal sets ---------------------------------------------------------------------------------------------
    class RefRobot(object):   # robot_kinematics.inverse_kinematics over the reference's compiled KDL
        def inverse_kinematics(self, position, orientation=None, seed=None):
            return chain.ref_ik(np.asarray(position, dtype=np.float64), np.asarray(orientation, dtype=np.float64),
                                np.asarray(seed, dtype=np.float64))[0]

    cfg.ROBOT = RefRobot()
    sc = S.make_scene(**SCENE_ARGS)
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        cfg.goal_set_proj = True
        cfg.use_standoff = case["use_standoff"]
        cfg.ik_parallel = case["ik_parallel"]
        cfg.increment_iks = bool(case.get("increment_iks", False))
        cfg.y_upsample = False
        cfg.scene_file = ""
        cfg.goal_idx = -1
        cfg.timesteps = 30
        cfg.get_global_param(30)
        env = H.make_ref_env(ns, sc, robot.body_points)
        target = env.objects[env.target_idx]
        target.pose = ns.util.pack_pose(target.pose_mat)
        target.compute_grasp = True
        target.attached = bool(case.get("attached", False))
        z_up = bool(case.get("z_upsample", False))
        target.seeds, target.grasp_potentials, target.grasp_vis_points = [], [], []
        for o in env.objects:
            o.compute_grasp = o is target
        rng = np.random.RandomState(case["seed"])
        pose_grasp = synthetic_grasps(rng, case["n_grasps"], target.pose_mat)
        traj = H.RefTrajectory(ns, np.zeros((30, 9)), S.START_CONF.copy(), S.START_CONF.copy(), goal_set=[], goal_idx=0)
        p = planner_mod.Planner.__new__(planner_mod.Planner)
        p.cfg, p.env, p.traj, p.lazy = cfg, env, traj, False
        p.cost = ns.cost.Cost(env)
        _print = builtins.print
        builtins.print = lambda *a, **k: None
        try:
            np.random.seed(5)      # (increment_iks draws its extra seeds from the global RNG)
            reach_raw, grasps_raw = p.solve_goal_set_ik(target, env, pose_grasp.copy(), z_upsample=z_up,
                                                        y_upsample=False, obj_coord=True)
            np.random.seed(5)
            p.solve_and_process_ik(target, pose_grasp.copy(), z_up)
            reach_proc, grasps_proc = np.array(target.reach_grasps), np.array(target.grasps)
            np.random.seed(7)
            p.setup_goal_set(env)
            reach_fin, grasps_fin = np.array(target.reach_grasps), np.array(target.grasps)
            pots_fin = np.array(target.grasp_potentials)
            p.grasp_init(env)
        finally:
            builtins.print = _print
        np.savez_compressed(
            os.path.join(out_dir, "goalset_%s.npz" % name), use_standoff=int(case["use_standoff"]),
            ik_parallel=int(case["ik_parallel"]), increment_iks=int(cfg.increment_iks), np_random_seed_ik=5,
            attached=int(target.attached), z_upsample=int(z_up), scene_args=np.array(repr(SCENE_ARGS)),
            sdf_checksum=np.float64(sc["sdf_grids"].astype(np.float64).sum()), body_points=robot.body_points,
            pose_grasp=pose_grasp, start=S.START_CONF, reach_raw=np.array(reach_raw), grasps_raw=np.array(grasps_raw),
            reach_processed=reach_proc, grasps_processed=grasps_proc, reach_final=reach_fin, grasps_final=grasps_fin,
            potentials_final=pots_fin, goal_idx=int(traj.goal_idx), end=np.array(traj.end), xi0=np.array(traj.data),
            np_random_seed=7)
        print(name, "raw", np.array(grasps_raw).shape, "processed", grasps_proc.shape, "final", grasps_fin.shape,
              "reach", reach_fin.shape, "goal_idx", traj.goal_idx)


if __name__ == "__main__":
    main()
