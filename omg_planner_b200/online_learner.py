"""Learner: the reference's online goal-set learner surface (omg/online_learner.py:61-259) over the fused
goal-scoring kernel.  Constructor arguments, method names and attributes follow the reference
(`Learner(env, traj, cost)`, `.update_goal()`, `.cost_vector()`, `.reset(traj)`, `.p`, `.t`), so
`Planner.plan`'s interleave (omg/planner.py:612-621) keeps working unchanged.

Split of work (BASELINE north_star: "online_learner goal reweighting stays on the host"):
  device  the expensive half of cost_vector -- interpolation to every goal, FK, SDF potentials, arc-length
          weighting and the sums -- ONE launch of omgb_goal_costs for all trajectories x goals
          (the reference: numpy FK of G*n' configurations + Cost.batch_obstacle_cost + two torch sums);
  host    the [B,G] vector algebra: the joint-difference "smooth" term, normalisation, FTL / FTC / Exp / MD / Proj
          (numpy, vectorised over the batch; the Bregman projection runs all trajectories and experts as rows of
          one masked iteration, row-for-row the reference's arithmetic).

Batched extension: a trajectory object whose .data is [B,n,9] with .goal_set [B,G,9], .goal_idx [B], .end [B,9]
and target.reach_grasps [B,G,c,9]; a plain [n,9] trajectory behaves exactly like the reference."""
import numpy as np


def _find_zero_rows(f, x0, x1, eps=1e-6, max_iter=100):
    """omg/online_learner.py:18-30, every row of the batch at once: a row stops moving at the iteration the
    reference would have returned."""
    x = (x0 + x1) / 2
    s = (x1 - x0) / 4
    done = np.zeros(x.shape[0], dtype=bool)
    for _ in range(max_iter):
        y = f(x)
        done |= np.abs(y) < eps
        if done.all():
            break
        move = ~done
        x = np.where(move, x - s * np.sign(y), x)
        s = np.where(move, s / 2, s)
    return x


def bregman_projection_rows(x, v, delta, w, max_iter=100, err=1e-6):
    """Bregman projection onto the simplex with the weighted, shifted entropy (omg/online_learner.py:32-58) for
    every row of x [R,G] / v [R,G] at once; rows leave the iteration when the reference's loop would break."""
    R = x.shape[0]
    alpha = np.zeros_like(x)
    y_out = np.zeros_like(x)
    active = np.ones(R, dtype=bool)
    target = 1 + np.sum(delta)
    for _ in range(max_iter):
        idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        a, xv, vv = alpha[idx], x[idx], v[idx]
        z = (a - vv) / w
        shifted = xv + delta
        lam = _find_zero_rows(lambda L: np.sum(shifted * np.exp(L[:, None] / w + z), axis=1) - target,
                              np.zeros(idx.size), np.max(w + vv, axis=1), err, max_iter)
        y_out[idx] = shifted * np.exp((lam[:, None] + a - vv) / w) - delta
        nxt = np.maximum(0, vv - lam[:, None] + w * np.log(delta / shifted))
        conv = np.sqrt(np.sum((a - nxt) ** 2, axis=1)) < err
        alpha[idx[~conv]] = nxt[~conv]
        active[idx[conv]] = False
    y = np.maximum(y_out, 0)
    return y / np.sum(y, axis=1, keepdims=True)


def _safe_div(a, b, eps=1e-8):   # omg/util.py:181-182
    return a / (b + eps)


class Learner(object):
    """An online learner that updates the goal distribution for the current trajectory (batch)."""

    def __init__(self, env, traj, cost):
        self.cfg = env.config
        self.env = env
        self.traj = traj
        self.cost = cost
        self._init_state(traj)
        target = self.env.objects[self.env.target_idx]
        if self.alg_name != "Proj" and len(target.reach_grasps) > 0:   # omg/online_learner.py:91-102
            costs = self.cost_vector()
            self._select(np.argmin(np.atleast_2d(costs), axis=-1))
            self.traj.interpolate_waypoints()

    # ---- state ------------------------------------------------------------------------------------
    def _init_state(self, traj):
        self.alg_name = self.cfg.ol_alg
        gs = np.asarray(traj.goal_set, dtype=np.float64)
        self.batched = np.asarray(traj.data).ndim == 3
        self.B = np.asarray(traj.data).shape[0] if self.batched else 1
        self.N = gs.shape[-2] if gs.ndim >= 2 else 0   # (no goals yet: the reference builds empty arrays too)
        self.T = self.cfg.optim_steps
        B, N = self.B, self.N
        self.Ti = np.zeros((B, N))
        self.Tis = []
        self.weights = np.ones(N)
        self.t = 0.0
        self._p = np.ones((B, N)) / max(N, 1)
        self.sum_costs = np.zeros((B, N))
        self.eta = np.sqrt(np.log(N + 1) / self.T)
        self.etas = [self.eta * (2 ** x) for x in [-2, -1, 0, 2, 4]]
        self.delta = np.ones(N) / (4 * N + 1)
        self.num_experts = len(self.etas)
        self._experts_p = np.ones((B, self.num_experts, N)) / max(N, 1)
        self.experts_costs = np.zeros((B, self.num_experts))
        self._q = np.ones((B, self.num_experts)) / self.num_experts

    @property
    def p(self):
        return self._p if self.batched else self._p[0]

    @property
    def q(self):
        return self._q if self.batched else self._q[0]

    @property
    def experts_p(self):
        return self._experts_p if self.batched else list(self._experts_p[0])

    def reset(self, traj):
        """omg/online_learner.py:251-263 (a full re-initialisation; the reference keeps the experts' state)."""
        self.traj = traj
        self._init_state(traj)

    # ---- views of the trajectory object ---------------------------------------------------------------
    def _goal_set(self):
        gs = np.asarray(self.traj.goal_set, dtype=np.float64)
        return gs if gs.ndim == 3 else gs[None]

    def _reach_goals(self):
        """The configurations scored for collision (omg/online_learner.py:123-127)."""
        if self.cfg.use_standoff:
            rg = np.asarray(self.env.objects[self.env.target_idx].reach_grasps, dtype=np.float64)
            rg = rg if rg.ndim == 4 else rg[None]
            return rg[:, :, -1, :]
        return self._goal_set()

    def _select(self, idx):
        idx = np.atleast_1d(np.asarray(idx)).astype(int)
        gs = self._goal_set()
        ends = gs[np.arange(self.B) % gs.shape[0], idx]
        if self.batched:
            self.traj.goal_idx = idx
            self.traj.end = ends
        else:
            self.traj.goal_idx = int(idx[0])
            self.traj.end = self.traj.goal_set[int(idx[0])]

    # ---- objective estimate ---------------------------------------------------------------------------------
    def cost_vector(self):
        cfg = self.cfg
        start = 1 + int((self.t / cfg.optim_steps) * cfg.timesteps) - 1     # omg/online_learner.py:108-110
        start = min(start, cfg.timesteps - 1)
        target = self.env.objects[self.env.target_idx]
        if cfg.traj_init == "grasp" and (
                len(target.reach_grasps) == 0
                or (cfg.use_standoff and np.asarray(target.reach_grasps).ndim == 2)):
            return np.zeros(1)
        data = np.asarray(self.traj.data, dtype=np.float64)
        data = data if self.batched else data[None]
        traj_start = data[:, start]                                          # [B,9]
        goal_set = self._goal_set()
        collision = self.cost.goal_costs(data, start, self._reach_goals())   # [B,G] fp32, one fused launch
        smooth = np.linalg.norm(np.diff(traj_start[:, None, :] - goal_set, axis=-1), axis=-1) ** 2   # sic (:151-153)
        potentials = cfg.base_obstacle_weight * collision + cfg.smoothness_base_weight * cfg.dist_eps * smooth
        if cfg.normalize_cost:
            potentials = potentials / np.linalg.norm(potentials, axis=-1, keepdims=True)
        return potentials if self.batched else potentials[0]

    # ---- update rules (omg/online_learner.py:162-235), vectorised over the batch ---------------------------
    def update_goal_dist(self):
        if self.alg_name == "Proj":
            self.Proj()
            return
        cv = np.atleast_2d(self.cost_vector())
        getattr(self, self.alg_name)(cv)

    def _one_hot(self, idx):
        self._p = np.zeros((self.B, self.N))
        self._p[np.arange(self.B), idx] = 1

    def FTL(self, cv):
        self.sum_costs = self.sum_costs + np.atleast_2d(cv)
        self.last_leader = np.argmin(self.sum_costs, axis=1)
        self._one_hot(self.last_leader)

    def FTC(self, cv):
        self.last_leader = np.argmin(np.atleast_2d(cv), axis=1)
        self._one_hot(self.last_leader)

    def Proj(self):
        data = np.asarray(self.traj.data, dtype=np.float64)
        data = data if self.batched else data[None]
        d = np.linalg.norm(data[:, -1][:, None, :] - self._goal_set(), axis=-1)
        self._one_hot(np.argsort(d, axis=1)[:, 0])

    def Exp(self, cv):
        cv = np.atleast_2d(cv)
        self.sum_costs = self.sum_costs + cv
        norm_sum = _safe_div(self.sum_costs, np.sum(self.sum_costs, axis=1, keepdims=True))
        p_new = np.exp(-self.eta * cv) * self._p
        self._p = p_new * 0.999 + norm_sum * 0.001
        self._p = _safe_div(self._p, np.sum(self._p, axis=1, keepdims=True))

    def MD(self, cv):
        cv = np.atleast_2d(cv)
        B, E, N = self.B, self.num_experts, self.N
        # every (trajectory, expert) projection is independent of the others: one masked row iteration
        v = (np.asarray(self.etas)[None, :, None] * cv[:, None, :]).reshape(B * E, N)
        p_all = bregman_projection_rows(self._experts_p.reshape(B * E, N), v, self.delta, self.weights).reshape(B, E, N)
        for i in range(E):
            p = p_all[:, i]
            self.experts_costs[:, i] = np.sum(cv * p, axis=1) + np.sum(self.weights * np.abs(p - self._experts_p[:, i]), axis=1)
            self._experts_p[:, i] = p
            # (sic) the mixture is re-weighted inside the expert loop, omg/online_learner.py:230-235
            self._q = self._q * np.exp(-1 * self.experts_costs)
            self._q = self._q / np.sum(self._q, axis=1, keepdims=True)
            self._p = np.sum(self._experts_p * self._q[:, :, None], axis=1)
            self._p = self._p / np.sum(self._p, axis=1, keepdims=True)

    def update_goal(self):
        """Take the argmax of the goal distribution (omg/online_learner.py:237-249).  Returns whether the goal
        changed (an array for a batched trajectory)."""
        self.t += 1
        self.update_goal_dist()
        old = np.atleast_1d(np.asarray(self.traj.goal_idx)).astype(int).copy()
        idx = np.argmax(self._p, axis=1)
        self._select(idx)
        self.Ti[np.arange(self.B), idx] += 1
        self.Tis.append(self.Ti if self.batched else self.Ti[0])
        changed = idx != old
        return changed if self.batched else bool(changed[0])


class DeviceLearnerState(object):
    """The Learner's state as device tensors plus the omgb_learner_update call (one launch for the whole batch):
    cost vector, FTL / FTC / Exp / MD / Proj, goal selection and the gather of the new goal rows, without leaving the
    GPU.  Built from a host Learner (so earlier host updates carry over) and written back with .store()."""

    def __init__(self, learner, device):
        import ctypes

        import torch

        from . import _lib

        self.learner, self.device, self._lib, self._ct, self._torch = learner, device, _lib, ctypes, torch
        cfg = learner.cfg
        B, G, E = learner.B, learner.N, learner.num_experts
        if G > 256:
            raise RuntimeError("the device learner handles at most 256 goals per trajectory")
        dev = lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)
        self.p, self.sum_costs = dev(learner._p), dev(learner.sum_costs)
        self.experts_p, self.experts_costs, self.q = dev(learner._experts_p), dev(learner.experts_costs), dev(learner._q)
        gs = np.asarray(learner.traj.goal_set, dtype=np.float64)
        self.shared = gs.ndim == 2 and B > 1
        self.goal_set = dev(gs if (gs.ndim == 3 or self.shared) else gs[None])
        self.reach = None
        self.c = 1
        if cfg.use_standoff:
            rg = np.asarray(learner.env.objects[learner.env.target_idx].reach_grasps, dtype=np.float64)
            self.reach = dev(rg if (rg.ndim == 4 or self.shared) else rg[None])
            self.c = rg.shape[-2]
            self.reach_goals = self.reach[..., -1, :].contiguous()
        else:
            self.reach_goals = self.goal_set
        self.goal_idx = dev(np.atleast_1d(np.asarray(learner.traj.goal_idx)).astype(np.int32), np.int32)
        if self.goal_idx.numel() != B:
            self.goal_idx = self.goal_idx.expand(B).contiguous()
        self.prm = _lib.LearnerParams()
        self.prm.alg = _lib.LEARNER_ALGS[learner.alg_name]
        self.prm.num_goals, self.prm.n_waypoints, self.prm.constraint_rows = G, cfg.timesteps, self.c
        self.prm.normalize_cost = int(bool(cfg.normalize_cost))
        self.prm.base_obstacle_weight = float(cfg.base_obstacle_weight)
        self.prm.smoothness_base_weight = float(cfg.smoothness_base_weight)
        self.prm.dist_eps = float(cfg.dist_eps)
        self.prm.eta = float(learner.eta)
        for k in range(E):
            self.prm.etas[k] = float(learner.etas[k])

    def update(self, engine, xi, end, goal_rows, done=None, selected=None, cost_vector=None):
        """Learner.update_goal for the batch: advances t, scores the goals (omgb_goal_costs) and updates / selects
        (omgb_learner_update); end [B,9] and goal_rows [B,c,9] are overwritten in place."""
        lrn, cfg, ct = self.learner, self.learner.cfg, self._ct
        lrn.t += 1
        first = min(1 + int((lrn.t / cfg.optim_steps) * cfg.timesteps) - 1, cfg.timesteps - 1)
        self.prm.first_waypoint = first
        coll = None
        if lrn.alg_name != "Proj":
            coll = engine.goal_costs(xi, first, self.reach_goals, float(cfg.time_interval), 0)
        vp = ct.c_void_p
        ptr = lambda t: None if t is None else vp(t.data_ptr())
        self._lib.check(self._lib.lib().omgb_learner_update(
            ct.byref(self.prm), xi.shape[0], ptr(xi), ptr(coll), ptr(self.goal_set), int(self.shared), ptr(self.reach),
            ptr(self.p), ptr(self.sum_costs), ptr(self.experts_p), ptr(self.experts_costs), ptr(self.q), ptr(done),
            ptr(self.goal_idx), ptr(end), ptr(goal_rows), ptr(cost_vector), ptr(selected),
            vp(self._torch.cuda.current_stream().cuda_stream)), "omgb_learner_update")

    def store(self):
        """Copy the state back into the host Learner (p, q, experts, sums) and the trajectory's goal."""
        lrn = self.learner
        lrn._p, lrn.sum_costs = self.p.cpu().numpy(), self.sum_costs.cpu().numpy()
        lrn._experts_p, lrn.experts_costs, lrn._q = (self.experts_p.cpu().numpy(), self.experts_costs.cpu().numpy(),
                                                       self.q.cpu().numpy())
        lrn._select(self.goal_idx.cpu().numpy())
