"""ChompEngine: thin host wrapper over the C ABI (include/omgb200.h).  torch is used only for device
memory and streams; every number on the hot path is produced by libomgb200.so."""
import ctypes

import numpy as np
import torch

from . import _lib
from .robot import PandaConstants

_vp = ctypes.c_void_p


def _hp(a):
    """host pointer of a contiguous numpy array (or None)"""
    return None if a is None else _vp(a.ctypes.data)


def _dp(t):
    return None if t is None else _vp(t.data_ptr())


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


class ChompEngine(object):
    def __init__(self, device=None, robot=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ChompEngine needs a CUDA device; the hot path has no CPU fallback")
        self.L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._h = _vp()
        _lib.check(self.L.omgb_scene_create(ctypes.byref(self._h), self.device.index), "omgb_scene_create")
        self._keep = {}
        self._metric_key = None
        self.num_objects = 0
        self.set_robot(robot if robot is not None else PandaConstants())

    def close(self):
        if self._h:
            self.L.omgb_scene_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene ------------------------------------------------------------------------------------
    def set_robot(self, robot, use_true_joint_origin=False):
        self.robot = robot
        pts = np.ascontiguousarray(robot.collision_points, dtype=np.float64)
        self.points_per_link = pts.shape[1]
        lo = np.ascontiguousarray(robot.joint_lower_limit, dtype=np.float64).reshape(-1)
        hi = np.ascontiguousarray(robot.joint_upper_limit, dtype=np.float64).reshape(-1)
        _lib.check(self.L.omgb_scene_set_robot(
            self._h, _hp(robot.pose_0), _hp(robot.tip2joint), _hp(robot.joint_axis), _hp(robot.joint_origin_true),
            int(use_true_joint_origin), _hp(robot.center_offset), _hp(pts), pts.shape[1], _hp(lo), _hp(hi)),
            "omgb_scene_set_robot")

    def set_sdf(self, sdf_grids, sdf_limits):
        """sdf_grids: torch CUDA fp32 [O,X,Y,Z] (env.sdf_torch, borrowed); sdf_limits: [O,10] (any)."""
        if not (sdf_grids.is_cuda and sdf_grids.dtype == torch.float32 and sdf_grids.is_contiguous()):
            raise RuntimeError("sdf_grids must be a contiguous fp32 CUDA tensor")
        lim = np.ascontiguousarray(sdf_limits.detach().cpu().numpy() if torch.is_tensor(sdf_limits) else sdf_limits,
                                   dtype=np.float32)
        o, x, y, z = sdf_grids.shape
        self._keep["grids"] = sdf_grids
        _lib.check(self.L.omgb_scene_set_sdf(self._h, _dp(sdf_grids), _hp(lim), o, x, y, z, _stream()),
                   "omgb_scene_set_sdf")
        self.num_objects = o

    def set_sdf_layout(self, layout):
        """0: the reference [O,X,Y,Z] layout (zero-copy); 1 / "quad": also keep the bricked quad copy (two 128-bit loads
        per trilinear sample; bit-identical results; 4x the grid bytes).  Call between set_sdf and set_objects."""
        code = {"plain": 0, "quad": 1}.get(layout, layout)
        _lib.check(self.L.omgb_scene_set_sdf_layout(self._h, int(code), _stream()), "omgb_scene_set_sdf_layout")

    def set_objects(self, pose_inv, epsilons, padding_scales, clearances, disables):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        pose_inv, epsilons, padding_scales, clearances, disables = map(
            f, (pose_inv, epsilons, padding_scales, clearances, disables))
        if pose_inv.shape != (self.num_objects, 4, 4):
            raise RuntimeError("pose_inv must be [O,4,4]")
        _lib.check(self.L.omgb_scene_set_objects(self._h, _hp(pose_inv), _hp(epsilons), _hp(padding_scales),
                                                 _hp(clearances), _hp(disables), _stream()), "omgb_scene_set_objects")

    def set_metric(self, cfg):
        """Upload cfg.Ainv and the goal-set projection; cached on (n, goal_set_proj, c, dt)."""
        c = cfg.constraint_rows
        ainv = np.ascontiguousarray(cfg.Ainv, dtype=np.float64)
        # keyed on the CONTENT of Ainv (a few KB, compared byte for byte: cheaper than hashing it on every call): an
        # array modified in place, or a new one at a recycled id(), must not leave a stale metric / projection on the
        # device
        key = (cfg.timesteps, bool(cfg.goal_set_proj), c, float(cfg.time_interval), ainv.shape, ainv.tobytes())
        if key == self._metric_key:
            return
        proj = cfg.projection_matrix(c)
        proj = None if proj is None else np.ascontiguousarray(proj, dtype=np.float64)
        _lib.check(self.L.omgb_scene_set_metric(self._h, cfg.timesteps, _hp(ainv), c, _hp(proj)), "omgb_scene_set_metric")
        self._metric_key = key

    def set_options(self, use_lower_bound=-1, use_longest_first=-1):
        """Toggle the exact accelerations (results are bit-identical either way)."""
        _lib.check(self.L.omgb_scene_set_options(self._h, int(use_lower_bound), int(use_longest_first)),
                   "omgb_scene_set_options")

    def load_scene(self, scene, cfg, sdf_layout=None):
        """Upload a scene dict (omg_planner_b200.scene.make_scene layout) with the per-object parameters
        Cost.compute_obstacle_cost_layer would build (omg/cost.py:303-328)."""
        from .cost import se3_inverse_f32

        grids = scene["sdf_grids"]
        grids = grids if torch.is_tensor(grids) else torch.from_numpy(grids)
        self.set_sdf(grids.to(self.device).contiguous(), scene["sdf_limits"])
        if sdf_layout is not None:
            self.set_sdf_layout(sdf_layout)
        num = len(scene["names"])
        poses = np.stack([se3_inverse_f32(scene["pose_mats"][i]) for i in range(num)])
        eps = np.full(num, cfg.epsilon, np.float32)
        clr = np.full(num, cfg.clearance, np.float32)
        pad = np.ones(num, np.float32)
        t = scene["target_idx"]
        eps[t], clr[t] = cfg.target_epsilon, cfg.target_clearance
        if scene.get("attached", False):
            clr[-1], eps[-1], pad[-1] = 0.0, 0.05, 0.5
        dis = np.array([1.0 if (nm == "floor" or nm in cfg.disable_collision_set) else 0.0
                        for nm in scene["names"]], np.float32)
        self.set_objects(poses, eps, pad, clr, dis)
        return self

    # ---- parameters -------------------------------------------------------------------------------
    @staticmethod
    def params_from(cfg, update=True, into=None, lsw_cache=None):
        """omgb_step_params_t snapshot of cfg.  `into`/`lsw_cache`: reuse a struct across calls (the per-call
        cost matters at ~0.1 ms per fused step)."""
        p = _lib.StepParams() if into is None else into
        p.n_waypoints = int(cfg.timesteps)
        p.goal_set_proj = int(bool(cfg.goal_set_proj))
        p.constraint_rows = int(cfg.constraint_rows)
        p.top_k_collision = int(cfg.top_k_collision)
        p.uncheck_finger_collision = int(cfg.uncheck_finger_collision)
        p.consider_finger = int(bool(cfg.consider_finger))
        p.allow_collision_point = int(cfg.allow_collision_point)
        p.pre_terminate = int(bool(cfg.pre_terminate))
        p.joint_limit_max_steps = int(cfg.joint_limit_max_steps)
        p.update = int(update)  # 0 info only, 1 always, 2 unless terminate
        p.time_interval = float(cfg.time_interval)
        p.obstacle_weight = float(cfg.obstacle_weight)
        p.smoothness_weight = float(cfg.smoothness_weight)
        p.step_size = float(cfg.step_size)
        p.clip_grad_scale = float(cfg.clip_grad_scale)
        p.terminate_smooth_loss = float(cfg.terminate_smooth_loss)
        lsw = cfg.link_smooth_weight
        if lsw_cache is not None and lsw_cache.get("obj") is lsw and lsw_cache.get("bytes") == getattr(lsw, "tobytes", bytes)():
            return p
        flat = np.asarray(lsw, dtype=np.float64).reshape(-1)
        for d in range(9):
            p.link_smooth_weight[d] = float(flat[d])
        if lsw_cache is not None and isinstance(lsw, np.ndarray):
            lsw_cache["obj"], lsw_cache["bytes"] = lsw, lsw.tobytes()
        return p

    # ---- hot path ---------------------------------------------------------------------------------
    def _check64(self, t, shape, name):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == tuple(shape)):
            raise RuntimeError("%s must be a contiguous fp64 CUDA tensor of shape %s" % (name, (shape,)))

    def step(self, cfg, xi, start, end, goal_rows=None, active=None, update=True, want_grad=False, debug=False,
             want_row_obs=False):
        """One fused CHOMP iteration in place on xi [B,n,9] (fp64 CUDA).  Returns info [B,16] (+grad,+dbg)."""
        self.set_metric(cfg)
        B, n = xi.shape[0], cfg.timesteps
        c = cfg.constraint_rows
        self._check64(xi, (B, n, 9), "xi"); self._check64(start, (B, 9), "start"); self._check64(end, (B, 9), "end")
        if c > 0:
            self._check64(goal_rows, (B, c, 9), "goal_rows")
        # (rows of inactive trajectories are never written: keep them defined)
        info = (torch.zeros if active is not None else torch.empty)((B, _lib.INFO_STRIDE), dtype=torch.float64,
                                                                    device=xi.device)
        grad = torch.empty_like(xi) if want_grad else None
        p = self.points_per_link
        dpot = torch.zeros((B, n, 10, p), dtype=torch.float32, device=xi.device) if debug else None
        dpts = torch.zeros((B, n, 10, p, 3), dtype=torch.float32, device=xi.device) if debug else None
        rows = torch.zeros((B, n), dtype=torch.float64, device=xi.device) if want_row_obs else None
        prm = self.params_from(cfg, update)
        _lib.check(self.L.omgb_chomp_step(self._h, ctypes.byref(prm), B, _dp(xi), _dp(start), _dp(end),
                                          _dp(goal_rows) if c > 0 else None, _dp(active), _dp(grad), _dp(info),
                                          _dp(dpot), _dp(dpts), _dp(rows), _stream()), "omgb_chomp_step")
        out = {"info": info}
        if want_row_obs:
            out["row_obs"] = rows
        if want_grad:
            out["grad"] = grad
        if debug:
            out["potentials"], out["points"] = dpot, dpts
        return out

    def plan(self, cfg, xi, start, end, goal_rows=None, iters=None, stop_on_terminate=False, first_step=1,
             history=False):
        """iters fused iterations with the reference's schedules (optimizer.py:63-80) in ONE persistent launch.
        history=True also records xi and the info row after every iteration (Planner.history_trajectories[1:] and
        Planner.info, omg/planner.py:621-622): out["hist_xi"] [iters,B,n,9], out["hist_info"] [iters,B,16];
        out["hist_all"] [iters+1,B,n,9] is the same storage with the initial trajectory in slot 0 (what
        Planner.history_trajectories holds), so the host side needs one copy and no concatenation."""
        self.set_metric(cfg)
        iters = cfg.optim_steps + cfg.extra_smooth_steps if iters is None else iters
        B, n, c = xi.shape[0], cfg.timesteps, cfg.constraint_rows
        self._check64(xi, (B, n, 9), "xi")
        sched = np.array([cfg.schedule(first_step + t) for t in range(iters)], dtype=np.float64).reshape(iters, 3)
        ow, sw, ss = (np.ascontiguousarray(sched[:, k]) for k in range(3))
        info = torch.empty((B, _lib.INFO_STRIDE), dtype=torch.float64, device=xi.device)
        done = torch.zeros((B,), dtype=torch.uint8, device=xi.device)
        hall = torch.empty((iters + 1, B, n, 9), dtype=torch.float64, device=xi.device) if history else None
        hx = None
        if history:
            hall[0].copy_(xi)
            hx = hall[1:]
        hi = torch.empty((iters, B, _lib.INFO_STRIDE), dtype=torch.float64, device=xi.device) if history else None
        prm = self.params_from(cfg, True)
        _lib.check(self.L.omgb_chomp_plan_history(self._h, ctypes.byref(prm), iters, _hp(ow), _hp(sw), _hp(ss),
                                                  int(stop_on_terminate), B, _dp(xi), _dp(start), _dp(end),
                                                  _dp(goal_rows) if c > 0 else None, _dp(done), _dp(info), _dp(hx),
                                                  _dp(hi), _stream()), "omgb_chomp_plan_history")
        out = {"info": info, "done": done}
        if history:
            out["hist_xi"], out["hist_info"], out["hist_all"] = hx, hi, hall
        return out

    def set_host_mode(self, mode):
        """0 auto (zero-copy on mapped pinned buffers, else pipelined staging), 1 staged, 2 staged + pipelined,
        3 zero-copy required (omgb_scene_set_host_mode)."""
        _lib.check(self.L.omgb_scene_set_host_mode(self._h, int(mode)), "omgb_scene_set_host_mode")

    def _host_ptr(self, a):
        """Data pointer of a numpy array, remembered per array object (ndarray.ctypes / __array_interface__ build a new
        Python object on every access: ~2 us each, five per call).  An array's buffer cannot move while the cache holds a
        reference to it (ndarray.resize refuses to reallocate a referenced array)."""
        cache = self._keep.setdefault("host_ptrs", {})
        e = cache.get(id(a))
        if e is not None and e[0] is a:
            return e[1]
        if len(cache) >= 16:
            cache.clear()
        p = a.ctypes.data
        cache[id(a)] = (a, p)
        return p

    def step_host(self, cfg, xi, start, end, goal_rows=None, info=None):
        """Reference-facing call with HOST numpy buffers (xi updated in place, info returned).  With pinned
        buffers (torch .pin_memory() / cudaHostRegister) the fused kernel reads and writes them directly over PCIe;
        pageable buffers go through pipelined staged copies.  `info`: optional [B,16] fp64 output buffer; by default
        a pinned buffer owned by the engine is reused (valid until the next call)."""
        self.set_metric(cfg)
        B, n, c = xi.shape[0], cfg.timesteps, cfg.constraint_rows
        for a, shp in ((xi, (B, n, 9)), (start, (B, 9)), (end, (B, 9))):
            if not (a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.shape == shp):
                raise RuntimeError("step_host buffers must be C-contiguous fp64 numpy arrays")
        if info is None:
            buf = self._keep.get("host_info")
            if buf is None or buf.shape[0] != B:
                buf = torch.empty((B, _lib.INFO_STRIDE), dtype=torch.float64).pin_memory()
                self._keep["host_info"] = buf
                self._keep["host_info_np"] = buf.numpy()
            info = self._keep["host_info_np"]
        if "host_prm" not in self._keep:
            self._keep["host_prm"], self._keep["host_lsw"] = _lib.StepParams(), {}
        prm = self.params_from(cfg, True, into=self._keep["host_prm"], lsw_cache=self._keep["host_lsw"])
        ptr = self._host_ptr
        _lib.check(self.L.omgb_chomp_step_host(self._h, ctypes.byref(prm), B, ptr(xi), ptr(start), ptr(end),
                                               ptr(goal_rows) if c > 0 else None, ptr(info),
                                               torch.cuda.current_stream().cuda_stream),
                   "omgb_chomp_step_host")
        return info

    def goal_costs(self, xi, first, goals, time_interval=0.1, uncheck_finger_collision=0):
        """Device half of Learner.cost_vector (omg/online_learner.py:104-150) for a batch: xi [B,n,9] fp64 CUDA,
        `first` = the waypoint the lines start from, goals [B,G,9] or [G,9] fp64 CUDA -> costs [B,G] fp32 CUDA."""
        B, n = xi.shape[0], xi.shape[1]
        self._check64(xi, (B, n, 9), "xi")
        shared = goals.dim() == 2
        G = goals.shape[-2]
        self._check64(goals, (G, 9) if shared else (B, G, 9), "goals")
        if not 0 <= first < n:
            raise RuntimeError("first waypoint out of range")
        costs = torch.empty((B, G), dtype=torch.float32, device=xi.device)
        _lib.check(self.L.omgb_goal_costs(self._h, B, _vp(xi.data_ptr() + 8 * 9 * first), n * 9, _dp(goals), G,
                                          int(shared), n - first, float(time_interval), int(uncheck_finger_collision),
                                          _dp(costs), _stream()), "omgb_goal_costs")
        return costs

    def batch_obstacle_cost(self, joints, arc_length=-1, start=None, time_interval=0.1,
                            uncheck_finger_collision=-1, want_grad=True):
        M = joints.shape[0]
        self._check64(joints, (M, 9), "joints")
        p = self.points_per_link
        pot = torch.empty((M, 10, p), dtype=torch.float32, device=joints.device)
        col = torch.empty((M, 10, p), dtype=torch.float32, device=joints.device)
        grad = torch.empty((M, 10, p, 3), dtype=torch.float32, device=joints.device) if want_grad else None
        _lib.check(self.L.omgb_batch_obstacle_cost(self._h, _dp(joints), M, int(arc_length), _dp(start),
                                                   float(time_interval), int(uncheck_finger_collision), _dp(pot),
                                                   _dp(grad), _dp(col), _stream()), "omgb_batch_obstacle_cost")
        return pot, grad, col


def sdf_loss_forward(pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales, clearances, disables):
    """Drop-in for omg_cuda.sdf_loss_forward (layers/omg_layers.cpp:24-49): same arguments, same outputs
    [potentials [N], potential_grads [N,3], collides [N]]; raises RuntimeError on non-CUDA / non-contiguous
    input like the reference's AT_ASSERT (omg_layers.cpp:5-7)."""
    args = (pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales, clearances, disables)
    for t in args:
        if not (torch.is_tensor(t) and t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError("sdf_loss_forward: every input must be a contiguous fp32 CUDA tensor")
    L = _lib.lib()
    n, o = points.shape[0], pose_init.shape[0]
    if sdf_grids.dim() != 4 or sdf_grids.shape[0] != o or tuple(sdf_limits.shape) != (o, 10):
        raise RuntimeError("sdf_loss_forward: shape mismatch")
    pot = torch.empty((n,), dtype=torch.float32, device=points.device)
    grad = torch.empty((n, 3), dtype=torch.float32, device=points.device)
    col = torch.empty((n,), dtype=torch.float32, device=points.device)
    ws = torch.empty((max(int(L.omgb_sdf_loss_workspace_bytes(o)), 16),), dtype=torch.uint8, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(L.omgb_sdf_loss(_dp(pose_init), _dp(sdf_grids), _dp(sdf_limits), _dp(points), _dp(epsilons),
                                   _dp(padding_scales), _dp(clearances), _dp(disables), n, o, sdf_grids.shape[1],
                                   sdf_grids.shape[2], sdf_grids.shape[3], _dp(pot), _dp(grad), _dp(col), _dp(ws),
                                   _stream()), "omgb_sdf_loss")
    return [pot, grad, col]
