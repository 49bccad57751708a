"""Shared test scaffolding: fake Env / Trajectory objects shaped like the reference's (omg/core.py:23-57,
243-257) so the parity tests read like reference usage, plus oracle drivers."""
import types

import numpy as np

from omg_planner_b200 import scene as S
from omg_planner_b200.config import ChompConfig
from omg_planner_b200.robot import PandaConstants


class FakeTrajectory(object):
    """omg/core.py:23-57 semantics."""

    def __init__(self, data, start, end, goal_set=None, goal_idx=0):
        self.data = np.array(data, dtype=np.float64)
        self.start = np.array(start, dtype=np.float64)
        self.end = np.array(end, dtype=np.float64)
        self.goal_set = goal_set if goal_set is not None else []
        self.goal_idx = goal_idx

    def set(self, new_traj):
        self.data = new_traj


def make_env(scene, cfg, robot=None, device="cuda"):
    """Object carrying what Cost/Optimizer read from omg.core.Env."""
    import torch

    robot = robot or PandaConstants()
    env = types.SimpleNamespace()
    env.config = cfg
    env.target_idx = scene["target_idx"]
    env.objects = []
    for i, name in enumerate(scene["names"]):
        env.objects.append(types.SimpleNamespace(name=name, pose_mat=np.array(scene["pose_mats"][i]), attached=False,
                                                 reach_grasps=[], grasps=[]))
    env.sdf_torch = torch.from_numpy(scene["sdf_grids"]).to(device)
    env.sdf_limits = torch.from_numpy(scene["sdf_limits"]).to(device)
    rk = types.SimpleNamespace(_pose_0=robot.pose_0, _tip2joint=robot.tip2joint, _joint_axis=robot.joint_axis,
                               _joint_origin=robot.joint_axis, center_offset=robot.center_offset)
    env.robot = types.SimpleNamespace(robot_kinematics=rk, collision_points=robot.collision_points,
                                      joint_lower_limit=robot.joint_lower_limit,
                                      joint_upper_limit=robot.joint_upper_limit)
    return env


MODES = {
    "fixed_topk": dict(goal_set_proj=False, use_standoff=True, top_k_collision=1000),
    "fixed_full": dict(goal_set_proj=False, use_standoff=True, top_k_collision=0),
    "goalset_standoff_topk": dict(goal_set_proj=True, use_standoff=True, top_k_collision=1000),
    "goalset_single_full": dict(goal_set_proj=True, use_standoff=False, top_k_collision=0),
    "goalset_standoff_topk200": dict(goal_set_proj=True, use_standoff=True, top_k_collision=200),
}


def mode_from_fixture(g):
    """mode dict of a tests/golden/chomp_*.npz fixture (tools/make_golden.py): [goal_set_proj, use_standoff,
    top_k_collision(, consider_finger)]."""
    m = [int(v) for v in g["mode"]]
    mode = dict(goal_set_proj=bool(m[0]), use_standoff=bool(m[1]), top_k_collision=m[2])
    if len(m) > 3 and m[3]:
        mode["consider_finger"] = True
    return mode


def goal_rows_for(mode, tails, ends):
    if not mode["goal_set_proj"]:
        return None
    return tails if mode["use_standoff"] else ends[:, None, :]


def engine_for(scene, cfg, robot=None):
    """ChompEngine with the scene uploaded the way Cost.sync does."""
    from omg_planner_b200.engine import ChompEngine

    return ChompEngine(robot=robot).load_scene(scene, cfg)


def oracle_steps(scene, mode, xi, start, end, rows, iters, robot_ref=None, cfg_kw=None):
    """Run the oracle for `iters` iterations on every trajectory; returns history [B,iters+1,n,9], infos."""
    from oracle import chomp_ref as R

    robot_ref = robot_ref or R.PandaRef()
    B, n = xi.shape[0], xi.shape[1]
    hist = np.zeros((B, iters + 1, n, 9)); infos = []
    for b in range(B):
        cfg = R.RefConfig(timesteps=n, **mode, **(cfg_kw or {}))
        opt = R.ChompRef(robot_ref, scene, cfg, xi[b], start[b], end[b], None if rows is None else rows[b])
        hist[b, 0] = xi[b]
        row = []
        for it in range(iters):
            row.append(opt.step())
            hist[b, it + 1] = opt.xi
        infos.append(row)
    return hist, infos
