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
  * cfg.increment_iks (omg/config.py:94, off by default): later grasp poses get extra IK seeds taken from earlier
    solutions -- ten drawn with np.random.choice after every group of four poses with the pool (planner.py:436-441), the
    solution whose hand is closest without it (:365-373).  The chains from the fixed seeds do not depend on earlier
    results and are still ONE launch; the extra-seed chains are one small launch per group / pose, in the reference's
    order, with the same RNG calls.
  * the simulated grasp files of a few YCB objects are filtered by ycb_special_case (omg/util.py:335-365)."""
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


PANDA_WRIST_LIMIT = 2.8973   # |q7| limit hard-coded in the reference's flip test (omg/planner.py:232-234)


def pitch_of(R):
    """The second angle of transforms3d.euler.mat2euler(R) (static xyz axes): atan2(-R[2,0], |(R[0,0], R[1,0])|)."""
    R = np.asarray(R, dtype=np.float64)
    return np.arctan2(-R[..., 2, 0], np.sqrt(R[..., 0, 0] * R[..., 0, 0] + R[..., 1, 0] * R[..., 1, 0]))


def ycb_special_case(pose_grasp, name):
    """omg/util.py:335-365: grasp-pose filters for the YCB objects the reference treats as edge cases.  Thin objects keep
    top-down grasps only; bowl / mug (/ the meat can, unreachable there: it is caught by the first branch) keep tilted
    ones, pushed 2 cm along the approach axis."""
    pose_grasp = np.asarray(pose_grasp)
    if name in ("037_scissors", "010_potted_meat_can", "061_foam_brick"):
        t = np.abs(pose_grasp[:, :3, 3])
        pose_grasp = pose_grasp[(t[:, 2] > 0.09) & (t[:, 1] > 0.02) & (t[:, 0] < 0.05)]
        if len(pose_grasp):
            pose_grasp = pose_grasp[np.abs(pitch_of(pose_grasp[:, :3, :3])) > 0.06]
    elif name in ("024_bowl", "025_mug"):
        angle = 50 if name == "024_bowl" else 30
        pose_grasp = pose_grasp[np.abs(pitch_of(pose_grasp[:, :3, :3])) > angle * np.pi / 180]
        push = np.eye(4)
        push[2, 3] = 0.02
        pose_grasp = np.matmul(pose_grasp, push)
    return pose_grasp


def spin_about_vertical(poses, pivot, bins=50):
    """Placement up-sampling (omg/planner.py:324-335): every pose rotated about the world z axis through `pivot` (the
    object's position) by `bins` angles over a full turn.  [P,4,4] -> [bins,4,4] by broadcasting, as the reference does
    (it is used with the single relative hand pose of an attached object)."""
    turns = np.stack([rotZ(a) for a in np.linspace(-np.pi, np.pi, bins)], axis=0)
    centred = np.array(poses, dtype=np.float64)
    centred[:, :3, 3] = centred[:, :3, 3] - pivot
    out = np.matmul(turns, centred)
    out[:, :3, 3] += pivot
    return out


def tilt_about_contact_line(poses, bins=10, depth=0.13):
    """omg/planner.py:337-348: every pose tilted about the hand's y axis through the finger contact point (`depth` along
    the approach axis) by `bins` angles in [-45, 45] degrees.  [P,4,4] -> [P*bins,4,4], pose-major."""
    tilts = np.stack([rotY(a) for a in np.linspace(-np.pi / 4, np.pi / 4, bins)], axis=0)[:, :3, :3]
    poses = np.asarray(poses, dtype=np.float64)
    reach = np.array([0, 0, depth])
    contact = poses[:, :3, :3].dot(reach) + poses[:, :3, 3]                     # [P,3]
    turned = np.matmul(poses[:, :3, :3], tilts[:, None])                          # [bins,P,3,3]
    origin = contact[None] - turned.dot(reach)                                    # [bins,P,3]
    out = np.tile(poses[:, None], (1, bins, 1, 1))
    out[:, :, :3, 3] = origin.transpose((1, 0, 2))
    out[:, :, :3, :3] = turned.transpose((1, 0, 2, 3))
    return out.reshape(-1, 4, 4)


def _pose_mat(obj):
    """4x4 object->world pose of an env object (omg/core.py:88-97 keeps both .pose_mat and the packed .pose)."""
    return np.asarray(obj.pose_mat, dtype=np.float64)


class GoalSetMixin(object):
    """Needs self.cfg, self.env, self.traj, self.cost (provided by Planner)."""

    # ---- the IK solver for the scene's robot ---------------------------------------------------------------
    def ik_solver(self):
        if getattr(self, "_ik", None) is None:
            robot = self.env.robot
            lo = np.asarray(robot.joint_lower_limit, dtype=np.float64).reshape(-1).copy()
            hi = np.asarray(robot.joint_upper_limit, dtype=np.float64).reshape(-1).copy()
            pad = float(self.cfg.soft_joint_limit_padding)
            if pad != 0.2:   # the reference's KDL solver pads the URDF limits by a hard-coded 0.2 (robot_pykdl.py:122-138)
                lo[:7] = (lo[:7] - pad) + 0.2
                hi[:7] = (hi[:7] + pad) - 0.2
            self._ik = IkSolver(robot.robot_kinematics._pose_0, lo, hi)
        return self._ik

    def flip_grasp(self, old_grasps):
        """omg/planner.py:224-236: the same hand pose with the wrist (joint 7) half a turn the other way -- towards zero --
        and the mask of the results that stay inside the padded wrist limit."""
        flipped = np.array(old_grasps, dtype=np.float64)
        wrist = flipped[..., 6]
        flipped[..., 6] = wrist - np.sign(wrist) * np.pi
        inside = np.abs(flipped[..., 6]) < PANDA_WRIST_LIMIT - self.cfg.soft_joint_limit_padding
        return flipped, inside

    def solve_goal_set_ik(self, target_obj, env, pose_grasp, one_trial=False, z_upsample=False, y_upsample=False,
                          obj_coord=True):
        """omg/planner.py:296-455 + solve_one_pose_ik (:16-87): returns (reach_goal_set, standoff_goal_set) lists."""
        cfg = self.cfg
        object_pose = _pose_mat(target_obj)
        init_seed = np.asarray(self.traj.start, dtype=np.float64).reshape(-1)[:7]
        tail = cfg.reach_tail_length
        seeds = init_seed[None, :]
        if not one_trial:
            seeds = np.concatenate([seeds, UTIL_ANCHOR_SEEDS[: cfg.ik_seed_num, :7]], axis=0)

        hand_poses = np.array(pose_grasp, dtype=np.float64)
        if obj_coord:
            hand_poses = np.matmul(object_pose, hand_poses)                     # gripper -> world
        if z_upsample:
            hand_poses = spin_about_vertical(hand_poses, object_pose[:3, 3])
        if y_upsample:
            hand_poses = tilt_about_contact_line(hand_poses)

        # the reach tail: the grasp pose pulled back along its approach axis, nearest first (:350-355)
        retreat = np.tile(np.eye(4), (tail, 1, 1, 1))
        if cfg.use_standoff:
            retreat[:, 0, 2, 3] = -cfg.standoff_dist * np.linspace(0, 1, tail, endpoint=False)
        tail_poses = np.matmul(hand_poses, retreat)                             # [tail, P, 4, 4]

        num = hand_poses.shape[0]
        solved_poses = num - 1 if cfg.ik_parallel else num                       # (sic) :419
        if solved_poses <= 0:
            return [], []
        if cfg.use_standoff:
            # the chain of one (pose, seed): the farthest standoff pose from the seed, then tail poses 0..tail-1 each
            # from the previous solution (:45-62)
            order = [tail - 1] + list(range(tail))
            chain = np.stack([tail_poses[k, :solved_poses] for k in order], axis=1)               # [P, tail+1, 4, 4]
        else:
            chain = hand_poses[:solved_poses, None]
        targets = poses_to_targets(chain)
        sols, solved = self.ik_solver().solve_chains(targets, seeds)             # every fixed-seed chain: one launch

        finger_joint = np.array([0.04, 0.04])
        finger_joints = np.tile(finger_joint, (tail, 1))
        T = chain.shape[1]

        def goals_of(sol_ps, solved_ps):
            """solve_one_pose_ik's result lists for ONE pose over a list of seeds (:33-86), seeds in order."""
            reach, standoff = [], []
            for s_ in range(sol_ps.shape[0]):
                if solved_ps[s_] != T:
                    continue
                if cfg.use_standoff:
                    iks = [sol_ps[s_, 1 + k] for k in range(tail)]
                    if not target_obj.attached:
                        iks = iks[::-1]
                    reach_traj = np.stack(iks)
                    if np.linalg.norm(np.diff(reach_traj, axis=0)) < 2:          # smooth (:71-73)
                        standoff_ = iks[0] if not target_obj.attached else iks[-1]
                        reach.append(np.concatenate([reach_traj, finger_joints], axis=-1))
                        standoff.append(np.concatenate([standoff_, finger_joint]))
                else:
                    goal_ik = sol_ps[s_, 0]
                    reach.append(np.concatenate([goal_ik, finger_joint]))
                    standoff.append(np.concatenate([goal_ik, finger_joint]))
            return reach, standoff

        reach_goal_set, standoff_goal_set = [], []
        if not getattr(cfg, "increment_iks", False):
            for p in range(solved_poses):        # result order of the reference: pose-major, seeds in order
                r_, s_ = goals_of(sols[p], solved[p])
                reach_goal_set.extend(r_)
                standoff_goal_set.extend(s_)
            return list(reach_goal_set), list(standoff_goal_set)

        # ---- cfg.increment_iks: extra seeds from earlier solutions ---------------------------------------------
        ik = self.ik_solver()
        if cfg.ik_parallel:
            group = 4                                                            # the pool's size (:404)
            extra = np.zeros((0, 7))
            for i in range(0, num, group):
                idx = list(range(i, min(i + group, num - 1)))
                more = ik.solve_chains(targets[idx], extra) if (len(idx) and len(extra)) else None
                for k, p in enumerate(idx):
                    r_, s_ = goals_of(sols[p], solved[p])
                    if more is not None:
                        r2, s2 = goals_of(more[0][k], more[1][k])
                        r_, s_ = r_ + r2, s_ + s2
                    reach_goal_set.extend(r_)
                    standoff_goal_set.extend(s_)
                # (:436-441: ten of the solutions so far, drawn WITH replacement from the global numpy RNG)
                pick = np.random.choice(np.arange(len(standoff_goal_set)), min(len(standoff_goal_set), 10))
                extra = np.array(standoff_goal_set).reshape(-1, 9)[pick, :7]
        else:
            hand_center = np.empty((0, 3))
            for p in range(num):
                r_, s_ = goals_of(sols[p], solved[p])
                if len(standoff_goal_set) > 0 and len(hand_center) > 0:          # (:365-373)
                    dists = np.linalg.norm(hand_poses[p, :3, 3] - hand_center, axis=-1)
                    closest = np.argsort(dists)[:1]
                    extra = np.array(standoff_goal_set)[closest, :7].reshape(-1, 7)
                    m_sols, m_solved = ik.solve_chains(targets[p:p + 1], extra)
                    r2, s2 = goals_of(m_sols[0], m_solved[0])
                    r_, s_ = r_ + r2, s_ + s2
                reach_goal_set.extend(r_)
                standoff_goal_set.extend(s_)
                if len(s_) > 0:
                    hand_center = np.concatenate([hand_center, np.tile(hand_poses[p, :3, 3], (len(s_), 1))], axis=0)
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
                    pose_grasp = ycb_special_case(pose_grasp, target_obj.name)          # (:484)
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
