"""Goal-set construction: the reference's `Planner.solve_goal_set_ik`, `.flip_grasp`,
`.solve_and_process_ik`, `.setup_goal_set`, `.load_grasp_set`, `.load_goal_from_external` (omg/planner.py:224-597) with
the inverse kinematics of every (grasp pose, seed) pair solved in ONE launch of omgb_ik_solve instead of a 4-process
pool of PyKDL solvers, the hand-frame filter through omgb_hand_poses and the collision filter through the fused
batch_obstacle_cost kernel.  Method names, arguments and the order of the results follow the reference; `Planner`
inherits this mixin.

Reference behaviours kept on purpose:
  * cfg.ik_parallel (default True): the pool loop `range(i, min(i + processes, num - 1))` never reaches the LAST
    grasp pose (omg/planner.py:419); with ik_parallel False every pose is solved.
  * setup_goal_set's diversity filter records indices of `goal_set[1:]` (off by one against goal_set) and then indexes
    goal_set with them (omg/planner.py:551-577).
  * sampling uses the global numpy RNG (np.random.choice, omg/planner.py:566).
Not supported: cfg.increment_iks (off by default) -- its seeds depend on earlier results, which serialises the batch."""
import os

import numpy as np

from .ik import IkSolver, poses_to_targets

# omg/util.py:19-35: joint-space anchor configurations used as IK seeds
UTIL_ANCHOR_SEEDS = np.array([
    [2.5, 0.23, -2.89, -1.69, 0.056, 1.46, -1.27, 0.04, 0.04],
    [2.8, 0.23, -2.89, -1.69, 0.056, 1.46, -1.27, 0.04, 0.04],
    [2, 0.23, -2.89, -1.69, 0.056, 1.46, -1.27, 0.04, 0.04],
    [2.5, 0.83, -2.89, -1.69, 0.056, 1.46, -1.27, 0.04, 0.04],
    [0.049, 1.22, -1.87, -0.67, 2.12, 0.99, -0.85, 0.04, 0.04],
    [-2.28, -0.43, 2.47, -1.35, 0.62, 2.28, -0.27, 0.04, 0.04],
    [-2.02, -1.29, 2.20, -0.83, 0.22, 1.18, 0.74, 0.04, 0.04],
    [-2.2, 0.03, -2.89, -1.69, 0.056, 1.46, -1.27, 0.04, 0.04],
    [-2.5, -0.71, -2.73, -0.82, -0.7, 0.62, -0.56, 0.04, 0.04],
    [-2, -0.71, -2.73, -0.82, -0.7, 0.62, -0.56, 0.04, 0.04],
    [-2.66, -0.55, 2.06, -1.77, 0.96, 1.77, -1.35, 0.04, 0.04],
    [1.51, -1.48, -1.12, -1.55, -1.57, 1.15, 0.24, 0.04, 0.04],
    [-2.61, -0.98, 2.26, -0.85, 0.61, 1.64, 0.23, 0.04, 0.04],
])


def rotZ(a):   # omg/util.py:38-47
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])


def rotY(a):   # omg/util.py:50-59
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]])


def unpack_pose(pose):
    """omg/util.py:115-120: (x, y, z, qw, qx, qy, qz) -> 4x4 (transforms3d's quat2mat formula)."""
    w, x, y, z = [float(v) for v in pose[3:]]
    nq = w * w + x * x + y * y + z * z
    out = np.eye(4)
    if nq > np.finfo(np.float64).eps:
        s = 2.0 / nq
        X, Y, Z = x * s, y * s, z * s
        wX, wY, wZ, xX, xY, xZ, yY, yZ, zZ = w * X, w * Y, w * Z, x * X, x * Y, x * Z, y * Y, y * Z, z * Z
        out[:3, :3] = [[1.0 - (yY + zZ), xY - wZ, xZ + wY], [xY + wZ, 1.0 - (xX + zZ), yZ - wX],
                       [xZ - wY, yZ + wX, 1.0 - (xX + yY)]]
    out[:3, 3] = pose[:3]
    return out


def _pose_mat(obj):
    """4x4 object->world pose of an env object (omg/core.py:88-97 keeps both .pose_mat and the packed .pose)."""
    return np.asarray(obj.pose_mat, dtype=np.float64)


class GoalSetMixin(object):
    """Needs self.cfg, self.env, self.traj, self.cost (provided by Planner)."""

    # ---- the IK solver for the scene's robot ---------------------------------------------------------------
    def ik_solver(self):
        if getattr(self, "_ik", None) is None:
            robot = self.env.robot
            self._ik = IkSolver(robot.robot_kinematics._pose_0, robot.joint_lower_limit, robot.joint_upper_limit)
        return self._ik

    def flip_grasp(self, old_grasps):
        """omg/planner.py:224-236: wrist flipped by pi in joint space."""
        grasps = np.array(old_grasps[:])
        neg_mask, pos_mask = (grasps[..., -3] < 0), (grasps[..., -3] > 0)
        grasps[neg_mask, -3] += np.pi
        grasps[pos_mask, -3] -= np.pi
        limits = (grasps[..., -3] < 2.8973 - self.cfg.soft_joint_limit_padding) * (
            grasps[..., -3] > -2.8973 + self.cfg.soft_joint_limit_padding)
        return grasps, limits

    def solve_goal_set_ik(self, target_obj, env, pose_grasp, one_trial=False, z_upsample=False, y_upsample=False,
                          obj_coord=True):
        """omg/planner.py:296-455 + solve_one_pose_ik (:16-87): returns (reach_goal_set, standoff_goal_set) lists."""
        cfg = self.cfg
        if getattr(cfg, "increment_iks", False):
            raise RuntimeError("cfg.increment_iks is not supported by the batched IK (seeds would depend on earlier "
                               "solutions)")
        object_pose = _pose_mat(target_obj)
        init_seed = np.asarray(self.traj.start, dtype=np.float64).reshape(-1)[:7]
        tail = cfg.reach_tail_length
        anchor_seeds = UTIL_ANCHOR_SEEDS[: cfg.ik_seed_num].copy()
        seeds = init_seed[None, :] if one_trial else np.concatenate([init_seed[None, :], anchor_seeds[:, :7]], axis=0)

        pose_grasp = np.array(pose_grasp, dtype=np.float64)
        pose_grasp_global = np.matmul(object_pose, pose_grasp) if obj_coord else pose_grasp
        if z_upsample:   # placement: rotate about the object's global z (:324-335)
            global_rot_z = np.stack([rotZ(a) for a in np.linspace(-np.pi, np.pi, 50)], axis=0)
            translation = object_pose[:3, 3]
            pose_grasp_global[:, :3, 3] = pose_grasp_global[:, :3, 3] - object_pose[:3, 3]
            pose_grasp_global = np.matmul(global_rot_z, pose_grasp_global)
            pose_grasp_global[:, :3, 3] += translation
        if y_upsample:   # tilt about the antipodal contact line (:337-348)
            bin_num = 10
            global_rot_y = np.stack([rotY(a) for a in np.linspace(-np.pi / 4, np.pi / 4, bin_num)], axis=0)
            finger_translation = pose_grasp_global[:, :3, :3].dot(np.array([0, 0, 0.13])) + pose_grasp_global[:, :3, 3]
            local_rotation = np.matmul(pose_grasp_global[:, :3, :3], global_rot_y[:, None, :3, :3])
            delta_translation = local_rotation.dot(np.array([0, 0, 0.13]))
            pose_grasp_global = np.tile(pose_grasp_global[:, None], (1, bin_num, 1, 1))
            pose_grasp_global[:, :, :3, 3] = (finger_translation[None] - delta_translation).transpose((1, 0, 2))
            pose_grasp_global[:, :, :3, :3] = local_rotation.transpose((1, 0, 2, 3))
            pose_grasp_global = pose_grasp_global.reshape(-1, 4, 4)

        pose_standoff = np.tile(np.eye(4), (tail, 1, 1, 1))
        if cfg.use_standoff:
            pose_standoff[:, 0, 2, 3] = -cfg.standoff_dist * np.linspace(0, 1, tail, endpoint=False)
        standoff_grasp_global = np.matmul(pose_grasp_global, pose_standoff)     # [tail, P, 4, 4]

        num = pose_grasp_global.shape[0]
        solved_poses = num - 1 if cfg.ik_parallel else num                       # (sic) :419
        if solved_poses <= 0:
            return [], []
        if cfg.use_standoff:
            # the chain of one (pose, seed): the farthest standoff pose from the seed, then tail poses 0..tail-1 each
            # from the previous solution (:45-62)
            order = [tail - 1] + list(range(tail))
            chain = np.stack([standoff_grasp_global[k, :solved_poses] for k in order], axis=1)   # [P, tail+1, 4, 4]
        else:
            chain = pose_grasp_global[:solved_poses, None]
        sols, solved = self.ik_solver().solve_chains(poses_to_targets(chain), seeds)

        finger_joint = np.array([0.04, 0.04])
        finger_joints = np.tile(finger_joint, (tail, 1))
        reach_goal_set, standoff_goal_set = [], []
        T = chain.shape[1]
        for p in range(solved_poses):            # result order of the reference: pose-major, seeds in order
            for s in range(seeds.shape[0]):
                if solved[p, s] != T:
                    continue
                if cfg.use_standoff:
                    iks = [sols[p, s, 1 + k] for k in range(tail)]
                    if not target_obj.attached:
                        iks = iks[::-1]
                    reach_traj = np.stack(iks)
                    if np.linalg.norm(np.diff(reach_traj, axis=0)) < 2:          # smooth (:71-73)
                        standoff_ = iks[0] if not target_obj.attached else iks[-1]
                        reach_goal_set.append(np.concatenate([reach_traj, finger_joints], axis=-1))
                        standoff_goal_set.append(np.concatenate([standoff_, finger_joint]))
                else:
                    goal_ik = sols[p, s, 0]
                    reach_goal_set.append(np.concatenate([goal_ik, finger_joint]))
                    standoff_goal_set.append(np.concatenate([goal_ik, finger_joint]))
        return list(reach_goal_set), list(standoff_goal_set)

    def solve_and_process_ik(self, target_obj, pose_grasp, z_upsample, obj_coord=True):
        """omg/planner.py:238-293: IK, wrist-flip augmentation, removal of goals that need a large hand rotation."""
        cfg, env = self.cfg, self.env
        target_obj.reach_grasps, target_obj.grasps = self.solve_goal_set_ik(
            target_obj, env, pose_grasp, z_upsample=z_upsample, y_upsample=cfg.y_upsample, obj_coord=obj_coord)
        target_obj.grasp_potentials = []
        if cfg.augment_flip_grasp and not target_obj.attached and len(target_obj.reach_grasps) > 0:
            flip_grasps, flip_mask = self.flip_grasp(target_obj.grasps)
            flip_reach, flip_reach_mask = self.flip_grasp(target_obj.reach_grasps)
            mask = flip_mask
            target_obj.reach_grasps.extend(list(flip_reach[mask]))
            target_obj.grasps.extend(list(flip_grasps[mask]))
        target_obj.reach_grasps = np.array(target_obj.reach_grasps)
        target_obj.grasps = np.array(target_obj.grasps)

        if cfg.remove_flip_grasp and len(target_obj.reach_grasps) > 0 and not target_obj.attached:
            ik = self.ik_solver()
            start_hand_pose = ik.hand_poses(np.asarray(self.traj.start, dtype=np.float64).reshape(1, -1))[0]
            if cfg.use_standoff:
                n = 5
                goals = np.array(target_obj.reach_grasps[:, -1])
                t = np.linspace(0, 1, n + 2)[1:-1]
                start = np.asarray(self.traj.start, dtype=np.float64).reshape(-1)
                # multi_interpolate_waypoints(start, goals, n, 9, "linear") (omg/util.py:261-290)
                interp = (t[None, :, None] * goals[:, None, :] + (1.0 - t)[None, :, None] * start[None, None, :])
                target_hand_pose = ik.hand_poses(interp.reshape(-1, 9)).reshape(-1, n, 4, 4)
            else:
                target_hand_pose = ik.hand_poses(np.array(target_obj.grasps))[:, None]
            R_diff = np.matmul(target_hand_pose[..., :3, :3], start_hand_pose[:3, :3].transpose(1, 0))
            angle = np.abs(np.arccos((np.trace(R_diff, axis1=2, axis2=3) - 1) / 2))
            angle = angle * 180 / np.pi
            rot_masks = angle > cfg.target_hand_filter_angle
            z = target_hand_pose[..., :3, 0] / np.linalg.norm(target_hand_pose[..., :3, 0], axis=-1, keepdims=True)
            downward_masks = z[:, :, -1] < -0.3
            masks = (rot_masks + downward_masks).sum(-1) > 0
            target_obj.reach_grasps = list(target_obj.reach_grasps[~masks])
            target_obj.grasps = list(target_obj.grasps[~masks])

    def load_grasp_set(self, env):
        """omg/planner.py:457-500: grasp poses of every object that wants grasps -> IK goal sets.  Poses come from
        target_obj.grasps_poses when set, else from data/grasps/simulated/<name>.npy (absent files are skipped)."""
        cfg = self.cfg
        for i, target_obj in enumerate(env.objects):
            if not (getattr(target_obj, "compute_grasp", False) and (i == env.target_idx or not self.lazy)):
                continue
            if not target_obj.attached:
                if len(getattr(target_obj, "grasps_poses", [])) == 0:
                    path = os.path.join(getattr(cfg, "robot_model_path", ""), "..", "grasps", "simulated",
                                        "{}.npy".format(target_obj.name))
                    if not os.path.exists(path):
                        continue
                    try:
                        pose_grasp = np.load(path, allow_pickle=True).item()["transforms"]
                    except Exception:
                        pose_grasp = np.load(path, allow_pickle=True, fix_imports=True,
                                             encoding="bytes").item()[b"transforms"]
                    pose_grasp = np.matmul(pose_grasp, np.array(rotZ(np.pi / 2)))     # flip x, y (:481-482)
                    target_obj.grasps_poses = pose_grasp
                else:
                    pose_grasp = target_obj.grasps_poses
                z_upsample = False
            else:   # placement
                rel = (np.asarray(target_obj.rel_hand_pose_mat, dtype=np.float64)
                       if getattr(target_obj, "rel_hand_pose_mat", None) is not None
                       else unpack_pose(np.asarray(target_obj.rel_hand_pose, dtype=np.float64)))
                pose_grasp = np.linalg.inv(rel)[None]
                z_upsample = cfg.z_upsample
            self.solve_and_process_ik(target_obj, pose_grasp, z_upsample)

    def load_goal_from_external(self, grasp_list):
        """omg/planner.py:176-185: grasp poses detected elsewhere, in world coordinates."""
        target_obj = self.env.objects[self.env.target_idx]
        self.solve_and_process_ik(target_obj, np.array(grasp_list), False, obj_coord=False)
        target_obj.compute_grasp = True
        self.setup_goal_set(self.env)

    def setup_goal_set(self, env, filter_collision=True, filter_diversity=True):
        """omg/planner.py:502-597: drop goals in collision, thin out near-duplicates, sample at most
        cfg.goal_set_max_num."""
        cfg = self.cfg
        for i, target_obj in enumerate(env.objects):
            goal_set = target_obj.grasps
            reach_goal_set = target_obj.reach_grasps
            if len(goal_set) > 0 and getattr(target_obj, "compute_grasp", False):
                potentials, _, vis_points, collide = self.cost.batch_obstacle_cost(
                    goal_set, special_check_id=i, uncheck_finger_collision=-1)
                collide = collide.sum(-1).sum(-1).detach().cpu().numpy()
                potentials = potentials.sum(dim=(-2, -1)).detach().cpu().numpy()
                ik_goal_num = len(goal_set)
                if filter_collision:
                    collision_free = (collide <= cfg.allow_collision_point).nonzero()
                    goal_set = [goal_set[idx] for idx in collision_free[0]]
                    try:
                        reach_goal_set = [reach_goal_set[idx] for idx in collision_free[0]]
                    except Exception:
                        pass
                    potentials = potentials[collision_free[0]]
                    vis_points = vis_points[collision_free[0]]
                sample = False
                num = len(goal_set)
                indexes = range(num)
                if filter_diversity:
                    if num > 0:
                        unique_grasps = [goal_set[0]]
                        indexes = []
                        for j, joint in enumerate(goal_set[1:]):
                            dists = np.linalg.norm(np.array(unique_grasps) - joint, axis=-1)
                            if np.amin(dists) < 0.5:
                                continue
                            unique_grasps.append(joint)
                            indexes.append(j)                      # (sic) index into goal_set[1:]
                        num = len(indexes)
                if num > 0:
                    sample = True
                    sample_goals = np.random.choice(indexes, min(num, cfg.goal_set_max_num), replace=False)
                    target_obj.grasps = [goal_set[int(idx)] for idx in sample_goals]
                    target_obj.reach_grasps = [reach_goal_set[int(idx)] for idx in sample_goals]
                    target_obj.seeds = list(getattr(target_obj, "seeds", [])) + target_obj.grasps
                    target_obj.reach_grasps = np.array(target_obj.reach_grasps)
                    target_obj.grasp_potentials.append(potentials[sample_goals])
                    if not hasattr(target_obj, "grasp_vis_points"):
                        target_obj.grasp_vis_points = []
                    target_obj.grasp_vis_points.append(vis_points[sample_goals])
                if not sample:
                    target_obj.grasps = []
                    target_obj.reach_grasps = []
                    target_obj.grasp_potentials = []
                    target_obj.grasp_vis_points = []
            target_obj.compute_grasp = False
