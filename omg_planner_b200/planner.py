"""Planner: the reference's outer loop (omg/planner.py:90-135, 187-222, 600-653) over the fused kernels.

  * fixed goal, or a goal set without goal switching (ol_alg "Baseline"/"Proj"): the whole plan -- every iteration of
    every trajectory, early exit on terminate, history_trajectories and the info list -- is ONE persistent launch
    (omgb_chomp_plan_history) plus one info-only launch for the trajectories that did not terminate.
  * goal set with the online learner: per iteration omgb_goal_costs -> omgb_learner_update -> omgb_chomp_plan_step on
    one stream, no host synchronisation until the plan is over (_plan_with_device_learner).  cfg.host_learner = True
    (or more than 256 goals) keeps BASELINE north_star's split instead: goal scoring and the CHOMP step on the device,
    the learner's [B,G] update on the host (_plan_with_learner).

cfg.timeout (the reference's wall-clock stop, omg/planner.py:629, 3 s by default) is honoured by the two loops that
launch per iteration: the host-learner loop checks it every iteration like the reference; the device-learner loop
checks it every 8 iterations against the DEVICE's progress (an event per check, the host never more than 16 iterations
ahead) and stops enqueuing.  The single persistent launch of a fixed-goal plan cannot be interrupted from the host (a
70-iteration plan of 1024 trajectories is ~6 ms of device time).

Same names and results as the reference: `Planner(env, traj)`, `.plan(traj) -> info list`, `.history_trajectories`,
`.info`, `.selected_goals`, `.cost`, `.optim`, `.learner`, `.grasp_init(env)`.  Goal sets come in through
env.objects[target].grasps / .reach_grasps (built by goal_set.py from grasp poses, or given).
Batched extension: traj.data [B,n,9]; then `.info` is a list (per trajectory) of info lists and
`.history_trajectories` a list (per trajectory) of [len,n,9] arrays, each cut where the reference's loop would have
stopped for that trajectory."""
import time

import numpy as np
import torch

from . import _lib
from .cost import Cost
from .goal_set import GoalSetMixin
from .online_learner import Learner
from .optimizer import Optimizer

_INFO_FLAGS = ("terminate", "violate_limit", "execute", "failure_terminate")


def info_from_row(cfg, r, n):
    """info dict (omg/cost.py:509-530) from one omgb info row; the fused plan does not keep per-iteration gradients."""
    return {
        "collision_pts": None, "obs": r[0], "smooth": r[1], "grasp": 0, "weighted_obs": None, "weighted_smooth": None,
        "weighted_smooth_grad": r[7], "weighted_obs_grad": r[6], "weighted_grasp_grad": 0, "weighted_grasp": 0,
        "failure_terminate": bool(r[11]), "cost": r[2], "grad": r[5], "terminate": bool(r[8]), "collide": r[3],
        "standoff_idx": n - cfg.reach_tail_length if cfg.use_standoff else n - 1, "reach": r[4],
        "execute": bool(r[10]), "violate_limit": bool(r[9]), "p_in": r[12], "text": [],
    }


class InfoList(object):
    """The info list of one trajectory of a batched plan (Planner.info[b]): a read-only sequence whose dicts are built
    from the recorded info rows when they are looked at (a 1024-trajectory x 70-iteration plan has 72k of them)."""

    def __init__(self, cfg, rows, n, extra=None):
        self._cfg, self._rows, self._n, self._extra, self._cache = cfg, rows, n, extra, {}

    def __len__(self):
        return self._rows.shape[0] + (1 if self._extra is not None else 0)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        if not 0 <= k < len(self):
            raise IndexError(k)
        if k not in self._cache:
            row = self._rows[k] if k < self._rows.shape[0] else self._extra
            self._cache[k] = info_from_row(self._cfg, row, self._n)
        return self._cache[k]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _to_host_fresh(t):
    """Device tensor -> numpy view of a FRESH pinned tensor (torch's caching host allocator recycles the block once
    the caller drops the arrays): histories handed to the user must not alias a buffer the next plan overwrites, and
    a pageable copy of a 150 MB history plus its concatenation cost more than the plan itself."""
    buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return buf.numpy()


def _to_host(t, cache, key):
    """Device tensor -> numpy through a cached pinned staging buffer (large histories: pageable copies are slow)."""
    buf = cache.get(key)
    if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
        buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        cache[key] = buf
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return buf.numpy()


class Planner(GoalSetMixin):
    def __init__(self, env, traj, lazy=False):
        self.cfg = env.config
        self.env = env
        self.traj = traj
        self.cost = Cost(env)
        self.optim = Optimizer(env, self.cost)
        self.lazy = lazy
        if self.cfg.goal_set_proj:
            self._load_goals(env)
        else:
            self.traj.interpolate_waypoints()
        self.history_trajectories = []
        self.info = []
        self.selected_goals = []

    def _load_goals(self, env):
        """omg/planner.py:103-114 / 137-147.  Objects that carry grasp poses (compute_grasp) get their goal sets built
        here -- batched IK, flip augmentation, collision / diversity filters (goal_set.py); objects whose .grasps /
        .reach_grasps are already set are used as they are."""
        if getattr(self.cfg, "use_external_grasp", False):
            self.load_goal_from_external(self.cfg.external_grasps)
        elif self.cfg.scene_file == "" or self.cfg.traj_init == "grasp":
            self.load_grasp_set(env)
            self.setup_goal_set(env)
        else:
            self.load_goal_from_scene()
        self.grasp_init(env)
        self.learner = Learner(env, self.traj, self.cost)

    def update(self, env, traj):
        """omg/planner.py:121-152: re-plan in a changed scene / for a new trajectory with the same Cost."""
        self.cfg = env.config
        self.env = env
        self.traj = traj
        self.cost.env = env
        self.cost.cfg = env.config
        if len(env.objects) > 0:
            self.cost.target_obj = env.objects[env.target_idx]
        self.optim = Optimizer(env, self.cost)
        if self.cfg.goal_set_proj:
            self._load_goals(env)
        else:
            self.traj.interpolate_waypoints()
        self.history_trajectories, self.info, self.selected_goals = [], [], []

    def load_goal_from_scene(self):
        """omg/planner.py:154-174: goals saved with a scene file (<scene_path>/<scene_file>.mat: 'goals',
        'reach_grasps', optionally 'grasp_qualities' / 'grasp_potentials'); standoff is not used with them unless
        cfg.force_standoff."""
        import os

        import scipy.io as sio

        cfg = self.cfg
        path = os.path.join(getattr(cfg, "scene_path", ""), cfg.scene_file + ".mat")
        if cfg.traj_init == "scene" and not hasattr(cfg, "force_standoff"):
            cfg.use_standoff = False
        if os.path.exists(path):
            scene = sio.loadmat(path)
            cfg.goal_set_max_num = len(scene["goals"])
            self.traj.goal_set = scene["goals"]
            self.env.objects[self.env.target_idx].reach_grasps = scene["reach_grasps"]
            if "grasp_qualities" in scene:
                self.traj.goal_quality = scene["grasp_qualities"][0]
                self.traj.goal_potentials = scene["grasp_potentials"][0]
            else:
                self.traj.goal_quality = np.zeros(cfg.goal_set_max_num)
                self.traj.goal_potentials = np.zeros(cfg.goal_set_max_num)

    def grasp_init(self, env=None):
        """omg/planner.py:187-222: goal set from the target's grasps, initial goal by cfg.goal_idx, trajectory
        re-initialised towards it.  Batched: goal_set [B,G,9] / start [B,9] select per trajectory."""
        cfg, traj = self.cfg, self.traj
        env = self.env if env is None else env
        if cfg.scene_file == "" or cfg.traj_init == "grasp":
            if len(env.objects) > 0:
                target = env.objects[env.target_idx]
                # unconditional, as the reference (planner.py:192-194): a target whose IK / collision filter left no
                # goals must leave an EMPTY goal set (plan() then returns "planning not run"), not the previous
                # target's goals -- PlanningScene reuses one Trajectory across update_planner()/reset()
                traj.goal_set = target.grasps
                traj.goal_potentials = getattr(target, "grasp_potentials", [])
                if cfg.goal_set_proj and cfg.use_standoff and len(target.reach_grasps) > 0:
                    traj.goal_set = np.asarray(target.reach_grasps)[..., -1, :]
        if len(traj.goal_set) == 0:
            return
        gs = np.asarray(traj.goal_set, dtype=np.float64)
        start = np.asarray(traj.start, dtype=np.float64)
        batched = gs.ndim == 3
        st = start[:, None, :] if start.ndim == 2 else start
        proj_dist = np.linalg.norm((st - gs) * cfg.link_smooth_weight, axis=-1)
        traj.goal_quality = np.ones(proj_dist.shape)
        if cfg.goal_idx >= 0:
            idx = np.full(proj_dist.shape[:-1], cfg.goal_idx, dtype=int)
        elif cfg.goal_idx == -1:
            pots = np.asarray(getattr(traj, "goal_potentials", 0.0))
            pots = pots if pots.size else 0.0
            costs = pots + cfg.dist_eps * proj_dist
            idx = np.argmin(costs, axis=-1) if batched else np.argmin(costs)   # (the reference's flat argmin)
        else:
            idx = np.zeros(proj_dist.shape[:-1], dtype=int)
        if cfg.ol_alg == "Proj":
            idx = np.argmin(proj_dist, axis=-1)
        if batched:
            traj.goal_idx = np.asarray(idx)
            traj.end = gs[np.arange(gs.shape[0]), traj.goal_idx]
        else:
            traj.goal_idx = int(np.asarray(idx).reshape(-1)[0])
            traj.end = traj.goal_set[traj.goal_idx]
        traj.interpolate_waypoints()

    # ---- the outer loop ---------------------------------------------------------------------------------
    def plan(self, traj):
        """omg/planner.py:600-653."""
        cfg = self.cfg
        batched = np.asarray(traj.data).ndim == 3
        self.history_trajectories = [np.copy(traj.data)]
        self.info = []
        self.selected_goals = []
        start_time_ = time.time()
        alg_switch = cfg.ol_alg != "Baseline" and cfg.ol_alg != "Proj"
        if cfg.goal_set_proj and len(traj.goal_set) == 0:
            return self.info                                    # "planning not run"
        iters = cfg.optim_steps + cfg.extra_smooth_steps
        if cfg.goal_set_proj and alg_switch:
            if self.learner.N <= 256 and not getattr(cfg, "host_learner", False):
                self._plan_with_device_learner(traj, iters)
            else:
                self._plan_with_learner(traj, iters, start_time_)
        else:
            self._plan_fused(traj, iters)
        plan_time = time.time() - start_time_
        for lst in (self.info if batched else [self.info]):
            lst[-1]["time"] = plan_time
        return self.info

    def _plan_with_device_learner(self, traj, iters):
        """Goal switching without leaving the GPU: per iteration omgb_goal_costs -> omgb_learner_update (cost vector,
        FTL / FTC / Exp / MD / Proj, goal selection, new goal rows) -> omgb_chomp_plan_step (frozen trajectories,
        history), all on one stream with no host synchronisation until the plan is over.  The loop over the iterations
        runs inside the library (omgb_chomp_plan_goalset): one call enqueues the whole plan."""
        from .online_learner import DeviceLearnerState

        cfg, cost, lrn = self.cfg, self.cost, self.learner
        cost.sync()
        ecfg = cost.engine_cfg()
        eng = cost.engine
        xi, start, end, rows, batched = cost._traj_tensors(traj)
        B, n, c = xi.shape[0], xi.shape[1], ecfg.constraint_rows
        st = DeviceLearnerState(lrn, xi.device)
        done = torch.zeros((B,), dtype=torch.uint8, device=xi.device)
        info = torch.empty((B, _lib.INFO_STRIDE), dtype=torch.float64, device=xi.device)
        hist_all = torch.empty((iters + 1, B, n, 9), dtype=torch.float64, device=xi.device)
        hist_all[0].copy_(xi)                          # history_trajectories[0] (planner.py:606)
        hist_xi = hist_all[1:]
        hist_info = torch.zeros((iters, B, _lib.INFO_STRIDE), dtype=torch.float64, device=xi.device)
        n_sel = min(iters, cfg.optim_steps)
        selected = torch.zeros((max(n_sel, 1), B), dtype=torch.int32, device=xi.device)
        import ctypes
        vp = ctypes.c_void_p
        step0, t0 = self.optim.step, lrn.t
        # Host half of the loop, ahead of time: Optimizer.update()'s schedule of every iteration (written back into cfg
        # like the reference, omg/optimizer.py:59-80) and the waypoint each learner update scores from
        # (omg/online_learner.py:108-110); the loop itself runs inside omgb_chomp_plan_goalset.
        sched = np.empty((iters, 3), dtype=np.float64)
        for t in range(iters):
            self.optim.update()
            sched[t] = (cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size)
        firsts = np.array([min(1 + int(((t0 + t + 1) / cfg.optim_steps) * cfg.timesteps) - 1, cfg.timesteps - 1)
                           for t in range(n_sel)], dtype=np.int32)
        ecfg = cost.engine_cfg()
        eng.set_metric(ecfg)
        prm = eng.params_from(ecfg, True)
        coll = torch.empty((B, lrn.N), dtype=torch.float32, device=xi.device) if lrn.alg_name != "Proj" else None
        buf = _lib.GoalsetPlanBuffers()
        for k, v in (("xi", xi), ("start", start), ("end", end), ("goal_rows", rows), ("done", done), ("info", info),
                     ("hist_xi", hist_xi), ("hist_info", hist_info), ("goal_set", st.goal_set), ("reach", st.reach),
                     ("reach_goals", st.reach_goals), ("p", st.p), ("sum_costs", st.sum_costs),
                     ("experts_p", st.experts_p), ("experts_costs", st.experts_costs), ("q", st.q),
                     ("goal_idx", st.goal_idx), ("selected", selected), ("collision", coll)):
            setattr(buf, k, None if v is None else v.data_ptr())
        buf.goals_shared = int(st.shared)
        ran_c = ctypes.c_int(0)
        _lib.check(eng.L.omgb_chomp_plan_goalset(
            eng._h, ctypes.byref(prm), ctypes.byref(st.prm), iters, n_sel, vp(sched.ctypes.data),
            vp(firsts.ctypes.data), B, ctypes.byref(buf), float(cfg.timeout) if cfg.timeout != -1 else -1.0,
            ctypes.byref(ran_c), vp(torch.cuda.current_stream().cuda_stream)), "omgb_chomp_plan_goalset")
        ran = ran_c.value
        lrn.t = t0 + min(ran, n_sel)
        if ran < iters:                                                  # stopped by cfg.timeout: cfg as the loop left it
            self.optim.step = step0 + ran - 1
            self.optim.update()
            iters, hist_all, hist_info = ran, hist_all[:ran + 1], hist_info[:ran]
            n_sel = min(n_sel, ran)
        stage = self.__dict__.setdefault("_stage", {})
        h_info, hist, sel = (_to_host(hist_info, stage, "info").copy(), _to_host_fresh(hist_all),
                             selected.cpu().numpy())
        term = h_info[:, :, 8] > 0
        term[0] = False
        stopped = term.any(0)
        stop = np.where(stopped, term.argmax(0), iters - 1)
        self.optim.step = step0 + int(stop.max()) + 1      # one Optimizer.update() per iteration the loop ran
        final = None
        if (~stopped).any():
            active = torch.from_numpy((~stopped).astype(np.uint8)).to(xi.device)
            self.optim.update()
            ecfg = cost.engine_cfg()
            final = eng.step(ecfg, xi, start, end, rows, active=active, update=0)["info"].cpu().numpy()
        st.store()
        self._assemble(traj, batched, hist, h_info, stop, stopped, final, xi.cpu().numpy(), sel=sel, n_sel=n_sel)
        # Learner.Ti: how often every goal was selected over the iterations the trajectory's own loop ran
        cut = np.minimum(stop + 1, n_sel)
        live = np.arange(sel.shape[0])[:, None] < cut[None, :]
        flat = (np.arange(B)[None, :] * lrn.N + sel)[live]
        lrn.Ti = np.bincount(flat, minlength=B * lrn.N).reshape(B, lrn.N).astype(np.float64)

    def _assemble(self, traj, batched, hist, h_info, stop, stopped, final, new_xi, sel=None, n_sel=0):
        """history_trajectories / info (/ selected_goals) per trajectory, each cut where the reference's loop would
        have stopped for it (omg/planner.py:627-635).  hist: [iters + 1, B, n, 9], the initial trajectory in slot 0.
        Batched plans get views and lazily built dicts."""
        cfg = self.cfg
        B, n = hist.shape[1], hist.shape[2]
        infos, hists, sels = [], [], []
        stop_l, stopped_l = np.asarray(stop).tolist(), np.asarray(stopped).tolist()
        sel_t = None if sel is None else np.ascontiguousarray(np.asarray(sel).T)    # [B, n_sel]: rows .tolist() in C
        for b in range(B):
            k = stop_l[b]
            # terminated: the state after the terminating iteration is dropped from the history (:634-635)
            hists.append(hist[:k + 1, b] if stopped_l[b] else hist[:k + 2, b])
            extra = None if stopped_l[b] else final[b]
            infos.append(InfoList(cfg, h_info[:k + 1, b], n, extra))
            if sel_t is not None:
                sels.append(sel_t[b, :min(k + 1, n_sel)].tolist())
        traj.set(new_xi if batched else new_xi[0])
        if batched:
            self.info, self.history_trajectories = infos, hists
            if sel is not None:
                self.selected_goals = sels
        else:
            self.info, self.history_trajectories = list(infos[0]), list(hists[0])
            if sel is not None:
                self.selected_goals = sels[0]
        return sels

    def _plan_with_learner(self, traj, iters, start_time_):
        """Goal switching with the learner's update on the host (BASELINE north_star's split; used for goal sets larger
        than the device learner's 256 goals, or with cfg.host_learner): the reference's loop verbatim over the fused
        Learner / Optimizer calls.  A batch runs until every trajectory has terminated; per-trajectory results are cut
        at the trajectory's own stop."""
        cfg = self.cfg
        batched = np.asarray(traj.data).ndim == 3
        B = np.asarray(traj.data).shape[0] if batched else 1
        infos, hist, sel = [], [np.copy(traj.data)], []
        stop = np.full(B, -1)
        for t in range(iters):
            if t < cfg.optim_steps:
                self.learner.update_goal()
                sel.append(np.copy(traj.goal_idx))
            info = self.optim.optimize(traj, force_update=True)
            infos.append(info)
            hist.append(np.copy(traj.data))
            term = (np.asarray(info.terminate) if hasattr(info, "terminate")
                    else np.array([i["terminate"] for i in (info if batched else [info])]))
            if t > 0:
                stop[(stop < 0) & term] = t
            if (stop >= 0).all():
                break
            if cfg.timeout != -1 and time.time() - start_time_ > cfg.timeout and t > 0:
                break
        final = None
        if (stop < 0).any():
            final = self.optim.optimize(traj, info_only=True)
        if not batched:
            self.info = infos + ([final] if final is not None else [])
            self.history_trajectories = hist if final is not None else hist[:-1]
            self.selected_goals = [int(s) for s in sel]
            return
        self.info, self.history_trajectories, self.selected_goals = [], [], []
        # a trajectory whose loop ended early keeps the state (and goal) it had then (the batch kept stepping it)
        data = np.array(traj.data)
        for b in np.nonzero(stop >= 0)[0]:
            data[b] = hist[stop[b] + 1][b]
            if sel:
                traj.goal_idx[b] = sel[min(stop[b], len(sel) - 1)][b]
                traj.end[b] = np.asarray(traj.goal_set)[b, traj.goal_idx[b]]
        traj.set(data)
        for b in range(B):
            last = stop[b] if stop[b] >= 0 else len(infos) - 1
            lst = [infos[t][b] for t in range(last + 1)]
            h = [hist[t][b] for t in range(last + 2)]
            if stop[b] >= 0:
                del h[-1]
            else:
                lst.append(final[b])
            self.info.append(lst)
            self.history_trajectories.append(np.stack(h))
            self.selected_goals.append([int(s[b]) for s in sel[:min(last + 1, cfg.optim_steps)]])

    def _plan_fused(self, traj, iters):
        """No goal switching: one persistent launch for the whole plan, one info-only launch afterwards."""
        cfg, cost = self.cfg, self.cost
        cost.sync()
        ecfg = cost.engine_cfg()
        xi, start, end, rows, batched = cost._traj_tensors(traj)
        B, n = xi.shape[0], xi.shape[1]
        first = self.optim.step + 1
        out = cost.engine.plan(ecfg, xi, start, end, rows, iters=iters, stop_on_terminate=True, first_step=first,
                               history=True)
        stage = self.__dict__.setdefault("_stage", {})
        hist_info = _to_host(out["hist_info"], stage, "info").copy()    # [iters,B,16]
        hist = _to_host_fresh(out["hist_all"])                          # [iters + 1,B,n,9]
        term = hist_info[:, :, 8] > 0
        term[0] = False                                                 # the t > 0 rule (planner.py:627)
        stopped = term.any(0)
        stop = np.where(stopped, term.argmax(0), iters - 1)             # last iteration run, per trajectory
        # the reference's Optimizer.step counter: one update() per optimize() call (optimizer.py:59-63)
        self.optim.step += int(stop.max()) + 1
        final = None
        if (~stopped).any():
            active = torch.from_numpy((~stopped).astype(np.uint8)).to(xi.device)
            self.optim.update()                                         # schedule of the info-only call
            ecfg = cost.engine_cfg()
            final = cost.engine.step(ecfg, xi, start, end, rows, active=active, update=0)["info"].cpu().numpy()
        self._assemble(traj, batched, hist, hist_info, stop, stopped, final, xi.cpu().numpy())
