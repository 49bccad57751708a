"""CPU: the host logic of goal-set construction (omg_planner_b200/goal_set.py: result ordering, the pool loop's
dropped last pose, wrist-flip augmentation, hand-rotation filter, collision / diversity filters, sampling) replayed
against fixtures recorded from the reference's own Planner methods with its own KDL
(tools/make_golden_goalset.py).  The three device calls the mixin makes are answered here by the CPU oracle (bit-exact
restatement of KDL; oracle Cost), so everything must match the fixtures exactly; the GPU versions of the same calls
are tested in tests/test_gpu_goal_set.py."""
import glob
import os
import types

import numpy as np
import pytest
import torch

import helpers as H
from omg_planner_b200 import scene as S
from omg_planner_b200.config import ChompConfig
from omg_planner_b200 import goal_set as GS
from omg_planner_b200.goal_set import GoalSetMixin
from omg_planner_b200.robot import PandaConstants
from oracle import chomp_ref as R
from oracle import kdl_ik_ref as K

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "goalset_*.npz")))


class OracleIk(object):
    def __init__(self, robot):
        self.chain = K.PandaChain(robot.pose_0, robot.joint_lower_limit, robot.joint_upper_limit)

    def solve_chains(self, targets, seeds):
        P, T, Sd = targets.shape[0], targets.shape[1], seeds.shape[0]
        sols, solved = np.zeros((P, Sd, T, 7)), np.zeros((P, Sd), np.int32)
        for p in range(P):
            for s in range(Sd):
                solved[p, s], sols[p, s] = self.chain.ik_chain(targets[p], seeds[s])
        return sols, solved

    def hand_poses(self, joints):
        return np.stack([self.chain.fk_hand(q) for q in np.asarray(joints).reshape(-1, np.asarray(joints).shape[-1])])


class OracleCost(object):
    def __init__(self, scene, cfg, body_points, attached=False):
        self.scene, self.robot = dict(scene, attached=attached), R.PandaRef(body_points=body_points)
        self.cfg = R.RefConfig(goal_set_proj=True, use_standoff=cfg.use_standoff)

    def batch_obstacle_cost(self, joints, special_check_id=0, uncheck_finger_collision=-1, **kw):
        pot, grad, col = R.batch_obstacle_cost(self.robot, self.scene, self.cfg, np.asarray(joints),
                                               uncheck_finger_collision=uncheck_finger_collision)
        return torch.from_numpy(pot), torch.from_numpy(grad), np.zeros(pot.shape + (12,)), torch.from_numpy(col)


def harness_targets(poses):
    """(position, quaternion xyzw) exactly as the fixture run produced them: util.pack_pose -> transforms3d.mat2quat,
    which tools/ref_harness.py stubs with scipy (transforms3d is not installed)."""
    from scipy.spatial.transform import Rotation

    poses = np.asarray(poses, dtype=np.float64)
    flat = poses.reshape(-1, 4, 4)
    out = np.zeros((flat.shape[0], 7))
    for i, T in enumerate(flat):
        x, y, z, w = Rotation.from_matrix(T[:3, :3]).as_quat()
        q = np.array([w, x, y, z])
        q = q if w >= 0 else -q
        out[i, :3] = T[:3, 3]
        out[i, 3:] = [q[1], q[2], q[3], q[0]]
    return out.reshape(poses.shape[:-2] + (7,))


def harness_object_pose(pose_mat):
    """unpack_pose(pack_pose(pose_mat)) (omg/planner.py:303) under the same stubs."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import ref_harness

    t = harness_targets(pose_mat)
    out = np.eye(4)
    out[:3, :3] = ref_harness._quat2mat([t[6], t[3], t[4], t[5]])
    out[:3, 3] = t[:3]
    return out


class HostPlanner(GoalSetMixin):
    def __init__(self, cfg, env, traj, cost, ik):
        self.cfg, self.env, self.traj, self.cost, self.lazy, self._ik = cfg, env, traj, cost, False, ik


def build(g, device_free=True):
    sc = S.make_scene(**eval(str(g["scene_args"])))
    assert abs(sc["sdf_grids"].astype(np.float64).sum() - float(g["sdf_checksum"])) < 1e-6
    cfg = ChompConfig(goal_set_proj=True, use_standoff=bool(g["use_standoff"]), ik_parallel=bool(g["ik_parallel"]),
                      goal_idx=-1, increment_iks=bool(int(g["increment_iks"])) if "increment_iks" in g.files else False)
    robot = PandaConstants(body_points=g["body_points"])
    env = types.SimpleNamespace(config=cfg, target_idx=sc["target_idx"], objects=[])
    for i, name in enumerate(sc["names"]):
        env.objects.append(types.SimpleNamespace(name=name, pose_mat=np.array(sc["pose_mats"][i]), attached=False,
                                                 reach_grasps=[], grasps=[], compute_grasp=i == sc["target_idx"],
                                                 grasp_potentials=[], grasp_vis_points=[], seeds=[]))
    rk = types.SimpleNamespace(_pose_0=robot.pose_0)
    env.robot = types.SimpleNamespace(robot_kinematics=rk, joint_lower_limit=robot.joint_lower_limit,
                                      joint_upper_limit=robot.joint_upper_limit)
    traj = types.SimpleNamespace(start=np.array(g["start"]), goal_set=[])
    env.objects[env.target_idx].attached = bool(int(g["attached"])) if "attached" in g.files else False
    return sc, cfg, robot, env, traj


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_goal_set_host_logic_matches_reference(path, monkeypatch):
    g = np.load(path)
    sc, cfg, robot, env, traj = build(g)
    target = env.objects[env.target_idx]
    z_up = bool(int(g["z_upsample"])) if "z_upsample" in g.files else False
    p = HostPlanner(cfg, env, traj, OracleCost(sc, cfg, g["body_points"], target.attached), OracleIk(robot))
    rng_seed = int(g["np_random_seed_ik"]) if "np_random_seed_ik" in g.files else 0
    # the product's own pose -> quaternion conversion: same goals, to the sensitivity of KDL's 1e-6 stop rule
    np.random.seed(rng_seed)
    reach, grasps = p.solve_goal_set_ik(target, env, g["pose_grasp"].copy(), z_upsample=z_up)
    assert np.array(grasps).shape == g["grasps_raw"].shape
    assert np.abs(np.array(grasps) - g["grasps_raw"]).max() < 1e-3   # (each is a solution to 1e-6 in task space)
    # with the fixture run's conversions everything is bit-identical
    monkeypatch.setattr(GS, "poses_to_targets", harness_targets)
    target.pose_mat = harness_object_pose(target.pose_mat)
    np.random.seed(rng_seed)
    reach, grasps = p.solve_goal_set_ik(target, env, g["pose_grasp"].copy(), z_upsample=z_up)
    np.testing.assert_array_equal(np.array(grasps), g["grasps_raw"])
    np.testing.assert_array_equal(np.array(reach), g["reach_raw"])
    np.random.seed(rng_seed)
    p.solve_and_process_ik(target, g["pose_grasp"].copy(), z_up)
    np.testing.assert_array_equal(np.array(target.grasps), g["grasps_processed"])
    np.testing.assert_array_equal(np.array(target.reach_grasps), g["reach_processed"])
    np.random.seed(int(g["np_random_seed"]))
    p.setup_goal_set(env)
    np.testing.assert_array_equal(np.array(target.grasps), g["grasps_final"])
    np.testing.assert_array_equal(np.array(target.reach_grasps), g["reach_final"])
    np.testing.assert_allclose(np.array(target.grasp_potentials), g["potentials_final"], rtol=1e-6, atol=1e-7)
    assert target.compute_grasp is False
