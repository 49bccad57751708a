"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy (fp64) restatement of ONE CHOMP iteration of the reference, one trajectory at a time, exactly as
the reference evaluates it (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this):

  kinematics            ycb_render/robotPose/robot_pykdl.py:148-215 (forward_kinematics_parallel,
                        incl. the joint-origin aliasing of :104) + omg/util.py:185-220 (wrap_*)
  body points/Jacobian  omg/cost.py:60-72, 92-110, 112-190
  SDF operator          oracle/sdf_loss_ref.c  (layers/sdf_matching_loss_kernel.cu) through
                        omg/cost.py:288-360 (per-object parameters)
  functional gradient   omg/cost.py:24-43
  obstacle term         omg/cost.py:362-423 (top-k branch with numpy duplicate-index '+=' semantics, and
                        the top_k_collision == 0 full-sum branch)
  smoothness term       omg/cost.py:425-449, omg/util.py:165-178, omg/config.py:199-220
  total / info          omg/cost.py:451-532
  CHOMP update          omg/optimizer.py:59-80, 88-113, 115-135, 148-174; omg/core.py:43-57

PARITY STATUS: the reference ships no tests/golden vectors for this path.  This file
is pinned against the reference's OWN Python, imported unmodified under stubs in the build container
(tools/ref_harness.py), on shared synthetic scenes: tools/make_golden.py writes tests/golden/*.npz from
the reference run and tests/test_oracle_golden.py replays them through this file.  The CUDA half of the
reference cannot be built (Eigen absent), so the SDF operator under both is oracle/sdf_loss_ref.c.
"""
import json
import os

import numpy as np

from . import sdf_loss_ref

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "omg_planner_b200", "data")


# ----------------------------------------------------------------------------------------------
# configuration (defaults of omg/config.py:29-104, 115)
# ----------------------------------------------------------------------------------------------
class RefConfig(object):
    def __init__(self, **kw):
        self.smoothness_base_weight = 0.1
        self.base_obstacle_weight = 1.0
        self.base_grasp_weight = 1.0
        self.cost_schedule_decay = 1
        self.cost_schedule_boost = 1.02
        self.base_step_size = 0.1
        self.step_decay_rate = 1.0
        self.joint_limit_max_steps = 10
        self.epsilon = 0.2
        self.target_epsilon = 0.1
        self.clearance = 0.01
        self.target_clearance = 0.0
        self.top_k_collision = 1000
        self.link_smooth_weight = np.ones(9)
        self.terminate_smooth_loss = 35
        self.goal_set_proj = True
        self.use_standoff = True
        self.pre_terminate = True
        self.uncheck_finger_collision = 0
        self.allow_collision_point = 5
        self.soft_joint_limit_padding = 0.2
        self.clip_grad_scale = 10.0
        self.disable_collision_set = []
        self.consider_finger = False
        self.reach_tail_length = 5
        self.timesteps = 30
        self.time_interval = 0.1
        # goal selection (omg/config.py:39,67,68,79)
        self.optim_steps = 50
        self.extra_smooth_steps = 20                      # omg/config.py:77
        self.ol_alg = "MD"
        self.dist_eps = 0.1
        self.normalize_cost = True
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)
        self.obstacle_weight = self.base_obstacle_weight
        self.smoothness_weight = self.smoothness_base_weight
        self.step_size = self.base_step_size
        self.set_timesteps(self.timesteps)

    def set_timesteps(self, steps):
        """omg/config.py:199-220 (get_global_param): dt keeps the total duration; K, A, Ainv."""
        self.time_interval = (0.1 * self.timesteps) / steps
        self.timesteps = steps
        self.diff_rule = np.array([[0, 0, -1, 1, 0, 0, 0], [0, 0, 1, -2, 1, 0, 0]], dtype=np.float64)
        self.K = [finite_difference_matrix(steps, self.diff_rule[r], self.time_interval, r + 1,
                                           with_end=not self.goal_set_proj) for r in range(2)]
        self.A = self.K[0].T.dot(self.K[0])
        self.Ainv = np.linalg.inv(self.A)


def finite_difference_matrix(n, rule, dt, order, with_end):
    """omg/util.py:165-178."""
    half = len(rule) // 2
    mat = np.zeros([n + 1, n])
    for i in range(n + 1):
        for j in range(-half, half):
            if 0 <= i + j < n:
                mat[i, i + j] = rule[j + half]
    if not with_end:
        mat[-1, -1] = 0
    return mat / (dt ** order)


# ----------------------------------------------------------------------------------------------
# robot
# ----------------------------------------------------------------------------------------------
class PandaRef(object):
    """Constants of robot_p3.pkl (robot_pykdl.py:98-112) + padded limits (omg/core.py:157-164)."""

    def __init__(self, body_points=None, soft_padding=0.2):
        with open(os.path.join(_DATA, "panda_constants.json")) as f:
            c = json.load(f)
        self.pose_0 = np.array(c["pose_0"])
        self.tip2joint = np.array(c["tip2joint"])
        self.joint_axis = np.array(c["joint_axis"])
        self.joint_origin = self.joint_axis  # sic: robot_pykdl.py:104 aliases origin to the axis list
        self.center_offset = np.array(c["center_offset"])
        lim = np.array(c["joint_limits"])
        self.lower = lim[None, :, 0].copy()
        self.upper = lim[None, :, 1].copy()
        self.lower[:, :-2] += soft_padding
        self.upper[:, :-2] -= soft_padding
        if body_points is None:
            with open(os.path.join(_DATA, "panda_body_points.json")) as f:
                body_points = json.load(f)["points"]
        self.body_points = np.array(body_points, dtype=np.float64)  # [10, p, 3]


def to_degrees_with_dummy(q):
    """omg/util.py:185-202 (wrap_value / wrap_values): rad -> deg, zero inserted at index 7."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    out = np.zeros([q.shape[0], q.shape[1] + 1])
    out[:, :7] = q[:, :7] / np.pi * 180
    out[:, 8:] = q[:, 7:] / np.pi * 180
    return out


_ANC = {  # omg/util.py:213-220 wrap_joint(j+1): ancestor joint ids of link j
    **{j: list(range(j + 1)) for j in range(7)}, 7: list(range(7)), 8: list(range(7)) + [8],
    9: list(range(7)) + [9]}
_COLS = {  # omg/util.py:205-210 wrap_index(j+1): DOF columns the link's gradient lands in
    **{j: list(range(j + 1)) for j in range(7)}, 7: list(range(7)), 8: list(range(8)),
    9: list(range(7)) + [8]}


def link_frames(robot, q_deg10):
    """robot_pykdl.py:148-215 with return_joint_info=True, offset=True.  q_deg10: [m,10] degrees."""
    m = q_deg10.shape[0]
    q = q_deg10 / 180.0 * np.pi
    rx_off = [0, -np.pi, np.pi, np.pi, -np.pi, np.pi, np.pi]
    out = np.zeros([m, 10, 4, 4])
    cur = np.eye(4)[None]
    for i in range(7):
        rz = np.tile(np.eye(4), [m, 1, 1])
        rz[:, 0, 0] = np.cos(q[:, i]); rz[:, 0, 1] = -np.sin(q[:, i])
        rz[:, 1, 0] = np.sin(q[:, i]); rz[:, 1, 1] = np.cos(q[:, i])
        c, s = np.cos(rx_off[i]), np.sin(rx_off[i])
        rx = np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1.0]])
        b = np.matmul(robot.pose_0[i][None], np.matmul(rz, rx))
        if i > 0:
            b[..., [1, 2]] *= -1
        cur = np.matmul(cur, b)
        out[:, i] = cur
    lf = np.tile(robot.pose_0[8], [m, 1, 1]); lf[:, 1, 3] += q[:, -2]
    rf = np.tile(robot.pose_0[9], [m, 1, 1]); rf[:, 1, 3] -= q[:, -1]
    out[:, 7] = np.matmul(out[:, 6], robot.pose_0[7])
    out[:, 8] = np.matmul(out[:, 7], lf)
    out[:, 9] = np.matmul(out[:, 7], rf)
    jp = np.matmul(out, robot.tip2joint)
    axes = np.matmul(jp[..., :3, :3], robot.joint_axis[..., None])[..., 0]
    origins = np.matmul(jp[..., :3, :3], robot.joint_origin[..., None])[..., 0] + jp[..., :3, 3]
    out = np.matmul(out, robot.center_offset)
    return out, origins, axes


def place_points(poses, pts):
    """omg/cost.py:60-72: poses [m,10,4,4], pts [10,p,3] -> [m,10,p,3]."""
    return np.einsum("mjab,jpb->mjpa", poses[..., :3, :3], pts) + poses[:, :, None, :3, 3]


# ----------------------------------------------------------------------------------------------
# obstacle operator
# ----------------------------------------------------------------------------------------------
def object_params(scene, cfg, attached=False):
    """omg/cost.py:303-328 (+ se3_inverse omg/util.py:129-135)."""
    num = len(scene["names"])
    poses = np.zeros((num, 4, 4), np.float32)
    eps = np.zeros(num, np.float32); pad = np.zeros(num, np.float32)
    clr = np.zeros(num, np.float32); dis = np.zeros(num, np.float32)
    for i, name in enumerate(scene["names"]):
        if name == "floor" or name in cfg.disable_collision_set:
            dis[i] = 1
        pad[i] = 1; eps[i] = cfg.epsilon; clr[i] = cfg.clearance
        rt = np.asarray(scene["pose_mats"][i], dtype=np.float64)
        inv = np.eye(4, dtype=np.float32)
        inv[:3, :3] = rt[:3, :3].T
        inv[:3, 3] = -1 * np.dot(rt[:3, :3].T, rt[:3, 3].reshape(3, 1)).reshape(3)
        poses[i] = inv
        if i == scene["target_idx"]:
            clr[i] = cfg.target_clearance; eps[i] = cfg.target_epsilon
    if attached:
        clr[-1] = 0.0; eps[-1] = 0.05; pad[-1] = 0.5
    return poses, eps, pad, clr, dis


def sdf_query(scene, cfg, x, uncheck_finger_collision, stats=None):
    """omg/cost.py:288-360: x [n,10,p,3] fp64 -> fp32 potentials [n,10,p], grads [n,10,p,3], collides."""
    n, m, p, _ = x.shape
    poses, eps, pad, clr, dis = object_params(scene, cfg, scene.get("attached", False))
    pts32 = x.astype(np.float32).reshape(-1, 3)  # omg/cost.py:136 (.float())
    pot, grad, col, pin = sdf_loss_ref.sdf_loss_forward(
        poses, scene["sdf_grids"], scene["sdf_limits"], pts32, eps, pad, clr, dis, return_pin=True)
    if stats is not None:
        stats["p_in"] = stats.get("p_in", 0) + pin
    pot = pot.reshape(n, m, p); grad = grad.reshape(n, m, p, 3); col = col.reshape(n, m, p)
    if uncheck_finger_collision == -1:  # omg/cost.py:350-353
        pot[:, -2:] *= 0.1; grad[:, -2:] *= 0.1; col[:, -2:] = 0
    return pot, grad, col


# ----------------------------------------------------------------------------------------------
# CHOMP cost terms
# ----------------------------------------------------------------------------------------------
def functional_gradient(v, a, jt, c, gc):
    """omg/cost.py:24-43.  v,a,gc [...,P,3]; c [...,P]; jt [...,P,J,3] -> cost [...], grad [...,P,J]."""
    speed = np.linalg.norm(v, axis=-1, keepdims=True)
    cost = np.sum(c * speed[..., 0], axis=-1)
    vhat = v / (speed + 1e-8)
    proj = np.eye(3) - vhat[..., :, None] * vhat[..., None, :]
    kappa = c[..., None, None] * (np.matmul(proj, a[..., None]) / (speed[..., None] ** 2 + 1e-8))
    pg = np.matmul(proj, gc[..., None])
    grad = np.sum(np.matmul(jt, speed[..., None] * pg - kappa), axis=-1)
    return cost, grad


def point_jacobians(origins, axes, x, j):
    """omg/cost.py:92-110 for link j: x [n,p,3] -> J^T [n,p,J,3] (linear part only)."""
    anc = _ANC[j]
    ax = axes[:, anc][:, None]       # [n,1,J,3]
    org = origins[:, anc][:, None]
    jt = np.cross(ax, x[:, :, None, :] - org)
    if j >= 8:  # prismatic finger joint: column is the axis itself (cost.py:106-108)
        jt[:, :, -1, :] = axes[:, anc[-1]][:, None]
    return jt


def time_derivatives(cfg, x, x_start, x_end):
    """omg/config.py:134-159 on omg/cost.py:168-173: x [n,...,3] -> velocity, acceleration."""
    n = x.shape[0]
    flat = x.reshape(n, -1)
    out = []
    for r in (1, 2):
        d = cfg.K[r - 1][: n + 1, :n].dot(flat)
        d[0] += cfg.diff_rule[r - 1][2] * x_start.reshape(-1) / (cfg.time_interval ** r)
        d[-2] += cfg.diff_rule[r - 1][4] * x_end.reshape(-1) / (cfg.time_interval ** r)
        d[-1] += cfg.diff_rule[r - 1][3] * x_end.reshape(-1) / (cfg.time_interval ** r)
        out.append(d[:-1].reshape(x.shape))
    return out


def obstacle_term(robot, scene, cfg, xi, start, end, stats=None):
    """omg/cost.py:362-423 (+112-190)."""
    n = xi.shape[0]
    poses, origins, axes = link_frames(robot, to_degrees_with_dummy(xi))
    x = place_points(poses, robot.body_points)                       # [n,10,p,3]
    pot, gpot, col = sdf_query(scene, cfg, x, cfg.uncheck_finger_collision, stats)
    x_s = place_points(link_frames(robot, to_degrees_with_dummy(start))[0], robot.body_points)[0]
    x_e = place_points(link_frames(robot, to_degrees_with_dummy(end))[0], robot.body_points)[0]
    v, a = time_derivatives(cfg, x, x_s, x_e)
    obs_grad = np.zeros_like(xi)
    obs_cost = np.zeros([n, 10])
    if stats is not None:
        stats["tie_slack"] = 0.0
    if cfg.top_k_collision == 0:
        for j in range(10):
            jt = point_jacobians(origins, axes, x[:, j], j)
            c_j, g_j = functional_gradient(v[:, j], a[:, j], jt, pot[:, j], gpot[:, j])
            obs_cost[:, j] += c_j
            obs_grad[:, _COLS[j]] += g_j.sum(1)
    else:
        order = np.argsort(pot.flatten())[-cfg.top_k_collision:]
        top_n, top_m, top_p = np.unravel_index(order, pot.shape)
        last = 10 if cfg.consider_finger else 8
        if stats is not None and len(order) == cfg.top_k_collision and (pot > 0).sum() > cfg.top_k_collision:
            # Which of several points tied at the k-th largest potential make the cut is decided by numpy's
            # unstable argsort (implementation-defined; observed to vary).  tie_slack = the most info["obs"] can
            # move with that choice: n * sum of c*|v| over the tied points of the links that are summed.
            tau = pot[top_n[0], top_m[0], top_p[0]]
            tn, tm, tp = np.nonzero(pot == tau)
            if len(tn) > 1:
                keep = tm < last
                stats["tie_slack"] = float(n * np.sum(pot[tn, tm, tp][keep] * np.linalg.norm(v[tn, tm, tp][keep], axis=-1)))
        for j in range(last):
            mask = top_m == j
            if not mask.any():
                continue
            sn, sp = top_n[mask], top_p[mask]
            jt = point_jacobians(origins, axes, x[:, j], j)[sn, sp]
            c_j, g_j = functional_gradient(v[sn, j, sp], a[sn, j, sp], jt, pot[sn, j, sp], gpot[sn, j, sp])
            obs_cost[:, j] += c_j                         # scalar added to every row (cost.py:416)
            cols = _COLS[j]
            rows = np.repeat(sn, len(cols)); cc = np.tile(cols, len(sn))
            obs_grad[rows, cc] += g_j.flatten()           # duplicate indices: last write wins (cost.py:421)
    return obs_cost, obs_grad, col.sum(), pot, gpot, x


def smooth_term(cfg, xi, start, end):
    """omg/cost.py:425-449."""
    w = np.asarray(cfg.link_smooth_weight)[None]
    ed = np.zeros([xi.shape[0] + 1, xi.shape[1]])
    ed[0] = cfg.diff_rule[0][2] * start / cfg.time_interval
    if not cfg.goal_set_proj:
        ed[-1] = cfg.diff_rule[0][3] * end / cfg.time_interval
    vel = cfg.K[0].dot(xi)
    loss = 0.5 * np.linalg.norm((vel + ed) * w, axis=1) ** 2
    grad = (cfg.A.dot(xi) + cfg.K[0].T.dot(ed)) * w
    return loss, grad


def total_cost(robot, scene, cfg, xi, start, end, goal, stats=None):
    """omg/cost.py:451-532.  goal: goal_set[goal_idx] (for goal_dist)."""
    s_loss, s_grad = smooth_term(cfg, xi, start, end)
    o_loss, o_grad, collide, pot, gpot, x = obstacle_term(robot, scene, cfg, xi, start, end, stats)
    s_sum, o_sum = s_loss.sum(), o_loss.sum()
    w_o_grad = np.clip(cfg.obstacle_weight * o_grad, -cfg.clip_grad_scale, cfg.clip_grad_scale)
    w_s_grad = cfg.smoothness_weight * s_grad
    cost = cfg.obstacle_weight * o_sum + cfg.smoothness_weight * s_sum
    grad = w_o_grad + w_s_grad
    goal_dist = np.linalg.norm(xi[-1] - goal) if cfg.goal_set_proj else 0
    terminate = bool((collide <= cfg.allow_collision_point) and cfg.pre_terminate and (goal_dist < 0.01)
                     and s_sum < cfg.terminate_smooth_loss)
    info = {
        "obs": o_sum, "smooth": s_sum, "cost": cost, "collide": collide, "reach": goal_dist,
        "weighted_obs": cfg.obstacle_weight * o_sum, "weighted_smooth": cfg.smoothness_weight * s_sum,
        "weighted_obs_grad": np.linalg.norm(w_o_grad), "weighted_smooth_grad": np.linalg.norm(w_s_grad),
        "grad": np.linalg.norm(grad), "gradient": grad, "terminate": terminate,
        "failure_terminate": bool((collide >= cfg.allow_collision_point * 10)
                                  or s_sum >= cfg.terminate_smooth_loss * 2.5),
        "execute": bool((collide <= cfg.allow_collision_point) and (s_sum < cfg.terminate_smooth_loss)),
        "cost_traj": cfg.obstacle_weight * o_loss.sum(-1) + cfg.smoothness_weight * s_loss[:-1],
        "standoff_idx": len(xi) - cfg.reach_tail_length if cfg.use_standoff else len(xi) - 1,
        "potentials": pot, "potential_grads": gpot, "points": x,
        "tie_slack": 0.0 if stats is None else stats.get("tie_slack", 0.0),
    }
    return cost, grad, info


# ----------------------------------------------------------------------------------------------
# optimizer
# ----------------------------------------------------------------------------------------------
def limit_violation(robot, curve):
    """omg/optimizer.py:137-146."""
    return (curve < robot.lower) * (robot.lower - curve) + (curve > robot.upper) * (robot.upper - curve)


def project_joint_limits(robot, cfg, curve):
    """omg/optimizer.py:148-164."""
    cnt = 0
    viol = limit_violation(robot, curve)
    while np.linalg.norm(viol) > 1e-2 and cnt < cfg.joint_limit_max_steps:
        vstar = cfg.Ainv.dot(viol)
        idx = np.unravel_index(np.abs(viol).argmax(), viol.shape)
        scale = np.abs(viol).max() / (np.abs(vstar[idx]) + 1e-8)
        curve = curve + scale * vstar
        viol = limit_violation(robot, curve)
        cnt += 1
    return curve


class ChompRef(object):
    """One trajectory's optimizer state; .step() == Optimizer.optimize(traj, force_update=True)
    (omg/optimizer.py:115-135) on a Trajectory (omg/core.py:23-57)."""

    def __init__(self, robot, scene, cfg, xi, start, end, goal_rows=None, goal=None):
        self.robot, self.scene, self.cfg = robot, scene, cfg
        # traj.goal_set[traj.goal_idx]; every assignment site keeps it equal to traj.end
        # (omg/planner.py:222, omg/online_learner.py:101,246)
        self.goal = np.array(end if goal is None else goal, dtype=np.float64)
        self.xi = np.array(xi, dtype=np.float64)
        self.start = np.array(start, dtype=np.float64)
        self.end = np.array(end, dtype=np.float64)
        # goal-set mode: rows the trajectory tail is projected onto: reach_grasps[goal_idx] [c,9]
        # with use_standoff, else goal_set[goal_idx][None] (omg/optimizer.py:93-98)
        self.goal_rows = None if goal_rows is None else np.atleast_2d(np.array(goal_rows, dtype=np.float64))
        self.iteration = 0
        self.stats = {}

    def schedule(self):
        """omg/optimizer.py:59-80."""
        self.iteration += 1
        c = self.cfg
        c.obstacle_weight = c.base_obstacle_weight * c.cost_schedule_decay ** self.iteration
        c.smoothness_weight = c.smoothness_base_weight * c.cost_schedule_boost ** self.iteration
        c.step_size = c.step_decay_rate ** self.iteration * c.base_step_size

    def step(self, info_only=False):
        cfg, robot = self.cfg, self.robot
        self.schedule()
        goal = self.goal if cfg.goal_set_proj else None
        pin0 = self.stats.get("p_in", 0)
        cost, grad, info = total_cost(robot, self.scene, cfg, self.xi, self.start, self.end, goal, self.stats)
        # in-bounds (body point, enabled object) pairs of this iteration -- the numerator of the roofline
        info["p_in"] = self.stats.get("p_in", 0) - pin0
        low = (self.xi < robot.lower - 5e-3).any()          # omg/optimizer.py:166-174 (sic)
        high = self.xi > robot.upper + 5e-3
        info["violate_limit"] = bool((low * high).any())
        info["terminate"] = bool(info["terminate"] and not info["violate_limit"])
        if info_only:
            return info
        if cfg.goal_set_proj:                                # omg/optimizer.py:88-113
            n = cfg.A.shape[0]
            c = self.goal_rows.shape[0]
            b = self.xi[-c:] - self.goal_rows
            sel = np.zeros([c, n]); sel[-c:, -c:] = np.eye(c)
            m = cfg.Ainv.dot(sel.T).dot(np.linalg.inv(sel.dot(cfg.Ainv.dot(sel.T))))
            update = (-cfg.step_size * cfg.Ainv.dot(grad) + cfg.step_size * m.dot(sel).dot(cfg.Ainv).dot(grad)
                      - m.dot(b))
        else:
            update = -cfg.step_size * cfg.Ainv.dot(grad)
        if cfg.consider_finger:                              # omg/core.py:43-51
            self.xi += update
        else:
            self.xi[:, :-2] += update[:, :-2]
        self.xi[:, -2:] = np.minimum(np.maximum(self.xi[:, -2:], 0), 0.04)
        self.xi = project_joint_limits(robot, cfg, self.xi)
        return info


def batch_obstacle_cost(robot, scene, cfg, joints, arc_length=-1, uncheck_finger_collision=-1, start=None):
    """omg/cost.py:192-286 (+ omg/config.py:162-187): potentials [M,10,p], grads [M,10,p,3], collides for M
    configurations; with arc_length > 0 the potentials are weighted by the fp32 workspace speed."""
    joints = np.asarray(joints, dtype=np.float64).reshape(-1, 9)
    x = place_points(link_frames(robot, to_degrees_with_dummy(joints))[0], robot.body_points)   # [M,10,p,3]
    pot, grad, col = sdf_query(scene, cfg, x, uncheck_finger_collision)
    if arc_length > 0:
        x32 = x.astype(np.float32).reshape(-1, arc_length, 10, x.shape[2], 3)                   # [G,n,10,p,3]
        xs = place_points(link_frames(robot, to_degrees_with_dummy(start))[0], robot.body_points)[0].astype(np.float32)
        inv_dt = np.float32(1.0 / cfg.time_interval)
        prev = np.concatenate([np.broadcast_to(xs[None, None], x32[:, :1].shape), x32[:, :-1]], axis=1)
        vel = x32 * inv_dt + prev * (-inv_dt)
        speed = np.sqrt((vel * vel).sum(-1)).reshape(pot.shape)
        pot = pot * speed
    return pot, grad, col
