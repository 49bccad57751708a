"""Hot-path configuration: the cfg fields Cost/Optimizer read (omg/config.py:29-104) and the derived
finite-difference matrices of get_global_param (omg/config.py:199-220, omg/util.py:165-178).

Not a re-implementation of the reference's flag system: a plain attribute bag the host mirror snapshots per
call (Optimizer.update writes schedules back into cfg, omg/optimizer.py:68-80)."""
import numpy as np


def get_diff_matrix(n, diff_rule, time_interval, order=1, with_end=True):
    """Rows i = 0..n of the order-th finite difference; omg/util.py:165-178."""
    half = len(diff_rule) // 2
    mat = np.zeros([n + 1, n])
    idx = np.arange(n + 1)
    for off in range(-half, half):
        col = idx + off
        ok = (col >= 0) & (col < n)
        mat[idx[ok], col[ok]] = diff_rule[off + half]
    if not with_end:
        mat[-1, -1] = 0
    return mat / (time_interval ** order)


class ChompConfig(object):
    _DEFAULTS = dict(
        smoothness_base_weight=0.1, base_obstacle_weight=1.0, base_grasp_weight=1.0, cost_schedule_decay=1,
        cost_schedule_boost=1.02, base_step_size=0.1, step_decay_rate=1.0, joint_limit_max_steps=10,
        optim_steps=50, extra_smooth_steps=20, epsilon=0.2, target_epsilon=0.1, clearance=0.01,
        target_clearance=0.0, top_k_collision=1000, terminate_smooth_loss=35, goal_set_proj=True,
        use_standoff=True, pre_terminate=True, uncheck_finger_collision=0, allow_collision_point=5,
        soft_joint_limit_padding=0.2, clip_grad_scale=10.0, consider_finger=False, reach_tail_length=5,
        timesteps=30, time_interval=0.1, report_cost=False, report_time=False, timeout=3.0,
        base_link="panda_link0", ol_alg="MD", dist_eps=0.1, normalize_cost=True, traj_init="grasp",
        # trajectory initialisation / outer loop (omg/config.py:63,89,96-99,69,129)
        traj_interpolate="cubic", dynamic_timestep=False, traj_delta=0.05, traj_max_step=50, traj_min_step=2,
        goal_idx=-2, silent=True, scene_file="",
        host_learner=False,   # True: keep the learner's [B,G] update on the host (Planner._plan_with_learner)
        # SDF assets (omg/config.py:54,55,60)
        target_size=1.0, obstacle_size=1, penalize_constant=5,
        # goal-set construction (omg/config.py:53,66,71,82,83,87,88,94,95,101,102)
        ik_clearance=0.03, goal_set_max_num=100, ik_seed_num=12, standoff_dist=0.08, remove_flip_grasp=True,
        augment_flip_grasp=True, target_hand_filter_angle=120, increment_iks=False, ik_parallel=True,
        y_upsample=False, z_upsample=True)

    def __init__(self, **kw):
        for k, v in self._DEFAULTS.items():
            setattr(self, k, v)
        self.link_smooth_weight = np.ones(9)
        self.disable_collision_set = []
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError("unknown config field %r" % k)
            setattr(self, k, v)
        self.obstacle_weight = self.base_obstacle_weight
        self.smoothness_weight = self.smoothness_base_weight
        self.grasp_weight = self.base_grasp_weight
        self.step_size = self.base_step_size
        self.get_global_param(self.timesteps)

    def get_global_param(self, steps=None):
        """omg/config.py:199-220 including the time_interval quirk (SURVEY A-18)."""
        steps = self.timesteps if steps is None else steps
        self.time_interval = (0.1 * self.timesteps) / steps
        self.timesteps = steps
        self.diff_rule_length = 7
        self.diff_rule = np.array([[0, 0, -1, 1, 0, 0, 0], [0, 0, 1, -2, 1, 0, 0], [0, -0.5, 1, 0, -1, 0.5, 0]])
        self.diff_matrices = [get_diff_matrix(steps, self.diff_rule[i], self.time_interval, i + 1,
                                              not self.goal_set_proj) for i in range(3)]
        self.A = self.diff_matrices[0].T.dot(self.diff_matrices[0])
        self.Ainv = np.linalg.inv(self.A)

    @property
    def constraint_rows(self):
        if not self.goal_set_proj:
            return 0
        return self.reach_tail_length if self.use_standoff else 1

    def projection_matrix(self, c=None):
        """M = Ainv C^T (C Ainv C^T)^-1 with C selecting the last c rows (omg/optimizer.py:102-107)."""
        c = self.constraint_rows if c is None else c
        if c == 0:
            return None
        n = self.A.shape[0]
        C = np.zeros([c, n])
        C[-c:, -c:] = np.eye(c)
        return self.Ainv.dot(C.T).dot(np.linalg.inv(C.dot(self.Ainv.dot(C.T))))

    def schedule(self, step):
        """Weights Optimizer.update() sets for iteration `step` (1-based; omg/optimizer.py:63-80)."""
        return (self.base_obstacle_weight * self.cost_schedule_decay ** step,
                self.smoothness_base_weight * self.cost_schedule_boost ** step,
                self.step_decay_rate ** step * self.base_step_size)
