"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's per-iteration goal scoring and goal-distribution update
(omg/online_learner.py), one trajectory at a time, on top of oracle/chomp_ref.py:

  cost_vector          omg/online_learner.py:104-160  (start index :109-113, linear interpolation to every goal
                       via omg/util.py:261-290, Cost.batch_obstacle_cost with arc_length omg/cost.py:192-286,
                       fp32 sums :147-150, the joint-axis np.diff "smooth" term :151-153, normalisation :159-160)
  find_zero / bp       omg/online_learner.py:18-58    (bisection; Bregman projection onto the simplex)
  FTL/FTC/Proj/Exp/MD  omg/online_learner.py:176-235
  update_goal          omg/online_learner.py:237-249
  the plan() interleave omg/planner.py:612-621 (update_goal, then Optimizer.optimize)

PARITY STATUS: pinned against the reference's own Learner imported unmodified under stubs in the build container
(tools/make_golden_learner.py -> tests/golden/learner_*.npz, replayed by tests/test_oracle_learner.py).  As for
chomp_ref.py the SDF operator under both is oracle/sdf_loss_ref.c.
"""
import numpy as np

from . import chomp_ref as R


def interpolate_to_goals(start, goals, n):
    """omg/util.py:261-290 with mode='linear': [G*n, 9], goal-major, the n interior points of linspace(0,1,n+2).
    scipy.interpolate.interp1d(kind='linear') is an unpinned dependency of the reference (requirements.txt);
    the scipy in this image (1.18, the one the fixtures were generated with) evaluates the convex combination
    ((t - x_lo)/(x_hi - x_lo)) * y_hi + ((x_hi - t)/(x_hi - x_lo)) * y_lo; older releases used
    slope*(t - x_lo) + y_lo, which differs by at most one ulp."""
    goals = np.asarray(goals, dtype=np.float64)
    t = np.linspace(0, 1, n + 2)[1:-1]
    w_hi, w_lo = (t - 0.0) / (1.0 - 0.0), (1.0 - t) / (1.0 - 0.0)
    return (w_hi[None, :, None] * goals[:, None, :] + w_lo[None, :, None] * start[None, None, :]).reshape(-1, start.shape[0])


def first_waypoint(t, optim_steps, timesteps):
    """omg/online_learner.py:108-110."""
    s = 1 + int((t / optim_steps) * timesteps) - 1
    return min(s, timesteps - 1)


def collision_costs(robot, scene, cfg, traj_start, goals, n):
    """The device half of cost_vector: sum over waypoints, links and body points of potential x workspace speed
    along the straight joint-space line from traj_start to every goal (fp32, omg/online_learner.py:134-150)."""
    q = interpolate_to_goals(traj_start, goals, n)
    pot, _, _ = R.batch_obstacle_cost(robot, scene, cfg, q, arc_length=n, uncheck_finger_collision=0, start=traj_start)
    pot = np.asarray(pot, dtype=np.float32)
    return pot.sum(axis=(-2, -1), dtype=np.float32).reshape(-1, n).sum(-1, dtype=np.float32)


def cost_vector(robot, scene, cfg, xi, goal_set, reach_goals, t, return_parts=False):
    """omg/online_learner.py:104-160.  goal_set: traj.goal_set [G,9]; reach_goals: the configurations scored for
    collision (reach_grasps[:, -1, :] with standoff, else goal_set)."""
    s = first_waypoint(t, cfg.optim_steps, cfg.timesteps)
    traj_start = np.asarray(xi[s], dtype=np.float64)
    n = cfg.timesteps - s
    coll = collision_costs(robot, scene, cfg, traj_start, reach_goals, n)
    smooth = np.linalg.norm(np.diff(traj_start - np.asarray(goal_set), axis=-1), axis=-1) ** 2   # sic: diff over joints
    pot = cfg.base_obstacle_weight * coll + cfg.smoothness_base_weight * cfg.dist_eps * smooth
    if cfg.normalize_cost:
        pot = pot / np.linalg.norm(pot)
    return (pot, coll, smooth) if return_parts else pot


def find_zero(f, x0, x1, eps=1e-6, max_iter=100):
    """omg/online_learner.py:18-30."""
    x, step = (x0 + x1) / 2, (x1 - x0) / 4
    for _ in range(max_iter):
        y = f(x)
        if abs(y) < eps:
            return x
        x -= step * np.sign(y)
        step /= 2
    return x


def bregman_projection(x, v, delta, w, max_iter=100, err=1e-6):
    """omg/online_learner.py:32-58."""
    alpha = np.zeros(len(x))
    for _ in range(max_iter):
        z = (alpha - v) / w
        target = 1 + np.sum(delta)
        shifted = x + delta
        lam = find_zero(lambda L: np.sum(shifted * np.exp(L / w + z)) - target, 0, np.max(w + v), err, max_iter)
        y = shifted * np.exp((lam + alpha - v) / w) - delta
        nxt = np.maximum(0, v - lam + w * np.log(delta / shifted))
        if np.linalg.norm(alpha - nxt, ord=2) < err:
            break
        alpha = nxt
    y = np.maximum(y, 0)
    return y / np.sum(y)


def _safe_div(dividend, divisor, eps=1e-8):
    """omg/util.py:181-182."""
    return dividend / (divisor + eps)


class LearnerRef(object):
    """State and updates of omg/online_learner.py:61-103, 162-259 for one trajectory."""

    def __init__(self, cfg, num_goals):
        self.cfg, self.N, self.T = cfg, num_goals, cfg.optim_steps
        self.alg = cfg.ol_alg
        self.t = 0.0
        self.p = np.ones(self.N) / self.N
        self.sum_costs = np.zeros(self.N)
        self.weights = np.ones(self.N)
        self.eta = np.sqrt(np.log(self.N + 1) / self.T)
        self.etas = [self.eta * (2 ** x) for x in [-2, -1, 0, 2, 4]]
        self.delta = np.ones(self.N) / (4 * self.N + 1)
        self.experts_p = [np.ones(self.N) / self.N for _ in self.etas]
        self.experts_costs = np.zeros(len(self.etas))
        self.q = np.ones(len(self.etas)) / len(self.etas)

    def update(self, cv, xi_last=None, goal_set=None):
        """update_goal_dist (omg/online_learner.py:162-235); returns argmax p."""
        if self.alg == "Proj":
            d = np.linalg.norm(xi_last - np.asarray(goal_set), axis=-1)
            self.p = np.zeros(self.N); self.p[np.argsort(d)[0]] = 1
        elif self.alg == "FTL":
            self.sum_costs = self.sum_costs + cv
            self.p = np.zeros(self.N); self.p[np.argmin(self.sum_costs)] = 1
        elif self.alg == "FTC":
            self.p = np.zeros(self.N); self.p[np.argmin(cv)] = 1
        elif self.alg == "Exp":
            self.sum_costs = self.sum_costs + cv
            norm_sum = _safe_div(self.sum_costs, np.sum(self.sum_costs))
            p_new = np.exp(-self.eta * cv) * self.p
            self.p = p_new * 0.999 + norm_sum * 0.001
            self.p = _safe_div(self.p, np.sum(self.p))
        elif self.alg == "MD":
            for i in range(len(self.etas)):
                p = bregman_projection(self.experts_p[i], self.etas[i] * cv, self.delta, self.weights)
                self.experts_costs[i] = np.dot(cv, p) + np.dot(self.weights, np.abs(p - self.experts_p[i]))
                self.experts_p[i] = p
                # (the reference re-weights q and re-mixes inside the expert loop, online_learner.py:230-235)
                self.q = self.q * np.exp(-1 * self.experts_costs)
                self.q = self.q / np.sum(self.q)
                self.p = sum(self.experts_p[k] * self.q[k] for k in range(len(self.etas)))
                self.p = self.p / np.sum(self.p)
        else:
            raise ValueError(self.alg)
        return int(np.argmax(self.p))
