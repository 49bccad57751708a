"""Cost: the reference's plugin surface for the obstacle/smoothness cost (omg/cost.py:12-532), with every
number produced by the fused sm_100a kernels.  Constructor arguments, public method names, argument
meaning and the info dict keys follow the reference so Planner/Learner keep working unchanged
(omg/planner.py:100-101,115; omg/online_learner.py:134).

Batched extension: a trajectory object whose .data is [B,n,9] (and .start/.end [B,9]) is processed in one
launch; a plain [n,9] trajectory behaves exactly like the reference."""
import numpy as np
import torch

from .engine import ChompEngine
from .sdf_matching_loss import SDFLoss


class _RobotView(object):
    """What ChompEngine.set_robot needs, read from the reference's Robot/robot_kinematics objects
    (omg/core.py:140-164; robot_pykdl.py:101-110)."""

    def __init__(self, robot):
        rk = robot.robot_kinematics
        f = lambda a: np.ascontiguousarray(np.array(a), dtype=np.float64)
        self.pose_0, self.tip2joint, self.joint_axis = f(rk._pose_0), f(rk._tip2joint), f(rk._joint_axis)
        # whatever the reference object holds as "origin" (aliased to the axis list at robot_pykdl.py:104)
        self.joint_origin_true = f(rk._joint_origin)
        self.center_offset = f(rk.center_offset)
        self.collision_points = f(robot.collision_points)
        self.joint_lower_limit, self.joint_upper_limit = f(robot.joint_lower_limit), f(robot.joint_upper_limit)


def se3_inverse_f32(rt):
    """World->object pose as fp32 (omg/util.py:129-135 semantics: fp64 product rounded to fp32)."""
    rt = np.asarray(rt, dtype=np.float64)
    out = np.eye(4, dtype=np.float32)
    out[:3, :3] = rt[:3, :3].T
    out[:3, 3] = -(rt[:3, :3].T @ rt[:3, 3])
    return out


class BatchInfos(object):
    """The per-trajectory info dicts of one batched Optimizer.optimize / Cost.compute_total_loss call: a read-only
    sequence over one [B,16] info array.  Dicts (omg/cost.py:509-530 keys) are built when indexed; the gradient
    [B,n,9] and the per-row obstacle costs stay on the device until the first dict is built.  `.rows` is the raw
    array (columns: omg_planner_b200._lib.INFO_KEYS), `.terminate` etc. are vectorised views for batch callers."""

    def __init__(self, cost, cfg, rows, n, grad_dev, rowobs_dev, xi_before, start, end):
        self._cost, self._cfg, self.rows, self._n = cost, cfg, rows, n
        self._grad_dev, self._rowobs_dev, self._host = grad_dev, rowobs_dev, None
        self._before, self._start, self._end = xi_before, start, end   # (xi_before: numpy, or a device tensor fetched on demand)
        self._cache = {}

    def __len__(self):
        return self.rows.shape[0]

    @property
    def terminate(self):
        return self.rows[:, 8] > 0

    @property
    def cost(self):
        return self.rows[:, 2]

    def gradient(self):
        """[B,n,9] numpy (one D2H copy, cached)."""
        if self._host is None:
            to_np = lambda t: t.cpu().numpy() if torch.is_tensor(t) else t
            self._host = (to_np(self._grad_dev), to_np(self._rowobs_dev))
        return self._host[0]

    def __getitem__(self, b):
        if isinstance(b, slice):
            return [self[i] for i in range(*b.indices(len(self)))]
        if b < 0:
            b += len(self)
        if not 0 <= b < len(self):
            raise IndexError(b)
        if b not in self._cache:
            grad = self.gradient()
            if torch.is_tensor(self._before):
                self._before = self._before.cpu().numpy()
            self._cache[b] = self._cost._info_dict(self._cfg, self.rows[b], self._n, grad[b], self._host[1][b],
                                                   self._before[b], self._start[b], self._end[b])
        return self._cache[b]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class Cost(object):
    def __init__(self, env):
        self.env = env
        self.cfg = env.config
        self.sdf_loss = SDFLoss()
        if len(self.env.objects) > 0:
            self.target_obj = self.env.objects[self.env.target_idx]
        self.engine = ChompEngine()
        self._stage = {}
        self._pinned = {}      # id(ndarray) -> (weakref to it, the pinned tensor it views): results handed out by evaluate()
        self._robot_sig = None
        self._sdf_sig = None
        self._obj_sig = None

    # ---- scene synchronisation (host) ----------------------------------------------------------------
    def object_params(self):
        """Per-object operator parameters (omg/cost.py:303-328)."""
        cfg, objs = self.cfg, self.env.objects
        num = len(objs)
        poses = np.zeros((num, 4, 4), np.float32)
        eps = np.full(num, cfg.epsilon, np.float32)
        pad = np.ones(num, np.float32)
        clr = np.full(num, cfg.clearance, np.float32)
        dis = np.zeros(num, np.float32)
        for i, ob in enumerate(objs):
            if ob.name == "floor" or ob.name in cfg.disable_collision_set:
                dis[i] = 1
            poses[i] = se3_inverse_f32(ob.pose_mat)
        t = self.env.target_idx
        clr[t], eps[t] = cfg.target_clearance, cfg.target_epsilon
        if getattr(objs[t], "attached", False):   # placing: table parameters (cost.py:325-328)
            clr[-1], eps[-1], pad[-1] = 0.0, 0.05, 0.5
        return poses, eps, pad, clr, dis

    def _scene_signature(self):
        """What the per-object operator parameters depend on, as one cheap comparable value: object poses, names,
        the target, its attached flag, and the cfg scalars of omg/cost.py:303-328.  Rebuilding and hashing the
        parameter arrays themselves on every call (round 1) cost more than the fused kernel for one trajectory."""
        cfg, objs = self.cfg, self.env.objects
        t = self.env.target_idx
        poses = np.array([ob.pose_mat for ob in objs], dtype=np.float64)
        return (poses.tobytes(), tuple(ob.name for ob in objs), t, bool(getattr(objs[t], "attached", False)),
                float(cfg.epsilon), float(cfg.clearance), float(cfg.target_epsilon), float(cfg.target_clearance),
                tuple(cfg.disable_collision_set))

    def sync(self):
        robot = self.env.robot
        pts = robot.collision_points
        sig = (id(pts), np.asarray(pts, dtype=np.float64).tobytes(),
               np.asarray(robot.joint_lower_limit).tobytes(), np.asarray(robot.joint_upper_limit).tobytes())
        if sig != self._robot_sig:
            self.engine.set_robot(_RobotView(robot), use_true_joint_origin=True)
            self._robot_sig = sig
        grids = self.env.sdf_torch
        ssig = (grids.data_ptr(), tuple(grids.shape), getattr(grids, "_version", 0))
        if ssig != self._sdf_sig:
            self.engine.set_sdf(grids, self.env.sdf_limits)
            self._sdf_sig, self._obj_sig = ssig, None
        osig = self._scene_signature()
        if osig != self._obj_sig:
            self.engine.set_objects(*self.object_params())
            self._obj_sig = osig

    # ---- reference API ----------------------------------------------------------------------------------
    def forward_poses(self, joints):
        """omg/cost.py:45-58: link poses, joint origins and joint axes of ONE configuration (degrees, with the dummy
        eighth joint) from the scene's robot_kinematics object -- the same delegation as the reference; the fused
        kernels do their own forward kinematics and never call this."""
        robot = self.env.robot
        poses, origins, axes = robot.robot_kinematics.forward_kinematics_parallel(
            joints[None, ...], base_link=self.cfg.base_link, return_joint_info=True)
        return poses[0], origins[0], axes[0]

    def forward_points(self, pose, pts, normals=None):
        """omg/cost.py:60-72: body points [m,3,p] through link poses [n,m,4,4] -> [p,n,m,3]."""
        r = pose[..., :3, :3]
        t = pose[..., :3, [3]]
        x = np.matmul(r, pts[None, ...]) + t
        if normals is None:
            return x.transpose([3, 1, 0, 2])
        normal = np.matmul(r, normals[None, ...])
        return np.concatenate([x, normal], 2).transpose([3, 1, 0, 2])

    def _staged(self, key, arr):
        """numpy fp64 -> device through a cached pinned staging buffer (async H2D on the current stream; the buffer
        is reused by the next call, which is ordered behind this copy on the same stream)."""
        dev = self.engine.device
        arr = np.asarray(arr, dtype=np.float64)
        slot = self._stage.get(key)
        if slot is None or slot[0].shape != arr.shape:
            slot = (torch.empty(arr.shape, dtype=torch.float64).pin_memory(),
                    torch.empty(arr.shape, dtype=torch.float64, device=dev))
            self._stage[key] = slot
        ent = self._pinned.get(id(arr))
        if ent is not None and ent[0]() is arr and ent[1].shape == slot[1].shape:
            # the array is the pinned result of the previous evaluate() (Optimizer.optimize put it back into
            # traj.data): copy straight from it, no staging memcpy
            slot[1].copy_(ent[1], non_blocking=True)
            return slot[1]
        np.copyto(slot[0].numpy(), arr)
        slot[1].copy_(slot[0], non_blocking=True)
        return slot[1]

    def _fresh_pinned_result(self, dev_tensor):
        """Device tensor -> a FRESH pinned host tensor (async copy on the current stream; torch's caching host allocator
        recycles the block once the caller drops the array).  Returned as (tensor, register) where register(ndarray)
        remembers the numpy view so that the next evaluate() can copy from it directly."""
        import weakref

        host = torch.empty(dev_tensor.shape, dtype=dev_tensor.dtype, pin_memory=True)
        host.copy_(dev_tensor, non_blocking=True)

        def register(arr):
            # entries of arrays the caller has dropped go first: their pinned blocks return to torch's host cache, so
            # a loop of optimize() calls cycles through two or three blocks instead of pinning new memory every call
            self._pinned = {k: v for k, v in self._pinned.items() if v[0]() is not None}
            self._pinned[id(arr)] = (weakref.ref(arr), host)
            return arr
        return host, register

    def _traj_tensors(self, traj):
        data = np.asarray(traj.data, dtype=np.float64)
        batched = data.ndim == 3
        xi = self._staged("xi", data if batched else data[None])
        B = xi.shape[0]
        bc = lambda a: np.broadcast_to(np.asarray(a, dtype=np.float64).reshape((-1, 9))[-B:] if np.ndim(a) > 1
                                       else np.asarray(a, dtype=np.float64)[None], (B, 9))
        start, end = self._staged("start", bc(traj.start)), self._staged("end", bc(traj.end))
        rows = None
        if self.cfg.goal_set_proj:
            c = self.engine_cfg().constraint_rows
            idx = np.atleast_1d(traj.goal_idx).astype(int)
            if self.cfg.use_standoff:   # omg/optimizer.py:93-98; batched: reach_grasps [B,G,c,9], goal_idx [B]
                rg = np.asarray(self.target_obj.reach_grasps)
                goal = rg[np.arange(B), idx] if rg.ndim == 4 else rg[idx]
            else:
                gs = np.asarray(traj.goal_set)
                goal = (gs[np.arange(B), idx] if gs.ndim == 3 else gs[idx])[:, None]
            rows = self._staged("rows", np.broadcast_to(goal, (B, c, 9)))
        return xi, start, end, rows, batched

    def engine_cfg(self):
        """Snapshot of cfg with the derived fields ChompEngine reads."""
        cfg = self.cfg
        if not hasattr(cfg, "constraint_rows"):
            class _View(object):
                pass
            v = _View()
            v.__dict__.update(dict(cfg) if isinstance(cfg, dict) else cfg.__dict__)
            v.constraint_rows = 0 if not cfg.goal_set_proj else (cfg.reach_tail_length if cfg.use_standoff else 1)
            ainv = np.asarray(cfg.Ainv)

            def projection_matrix(c, ainv=ainv):
                if c == 0:
                    return None
                n = ainv.shape[0]
                C = np.zeros([c, n]); C[-c:, -c:] = np.eye(c)
                return ainv.dot(C.T).dot(np.linalg.inv(C.dot(ainv.dot(C.T))))
            v.projection_matrix = projection_matrix
            return v
        return cfg

    def _info_dict(self, cfg, r, n, grad, rows, xi_before, start, end):
        """One info dict (omg/cost.py:509-530) from one omgb info row."""
        smooth_rows = self._smooth_rows(cfg, xi_before, start, end)
        return {
            "collision_pts": None, "obs": r[0], "smooth": r[1], "grasp": 0, "weighted_obs": cfg.obstacle_weight * r[0],
            "weighted_smooth": cfg.smoothness_weight * r[1], "weighted_smooth_grad": r[7], "weighted_obs_grad": r[6],
            "weighted_grasp_grad": 0, "weighted_grasp": 0, "gradient": grad, "failure_terminate": bool(r[11]),
            "cost": r[2], "grad": r[5], "terminate": bool(r[8]), "collide": r[3],
            "standoff_idx": n - cfg.reach_tail_length if cfg.use_standoff else n - 1, "reach": r[4],
            "execute": bool(r[10]), "violate_limit": bool(r[9]),
            "cost_traj": cfg.obstacle_weight * rows + cfg.smoothness_weight * smooth_rows[:-1],
            "p_in": r[12], "text": [],
        }

    @staticmethod
    def _smooth_rows(cfg, xi, start, end):
        """Per-row smoothness loss (omg/cost.py:429-445) on the host, only for info['cost_traj']."""
        w = np.asarray(cfg.link_smooth_weight)[None]
        vel = np.asarray(cfg.diff_matrices[0]).dot(xi)
        vel[0] -= start / cfg.time_interval
        if not cfg.goal_set_proj:
            vel[-1] += end / cfg.time_interval
        return 0.5 * np.linalg.norm(vel * w, axis=1) ** 2

    def evaluate(self, traj, update_mode=0):
        """One fused iteration; update_mode as omgb_step_params_t.update.  Returns (infos, new_xi, batched).
        One trajectory (the reference's shape): a plain info dict, everything on the host when the call returns.
        A batch: `infos` is a BatchInfos -- one [B,16] info array copied back with xi; the per-trajectory dicts, the
        gradient [B,n,9] and cost_traj are fetched from the device only if somebody looks at them."""
        self.sync()
        cfg = self.engine_cfg()
        xi, start, end, rows, batched = self._traj_tensors(traj)
        want_dbg = bool(getattr(self.cfg, "vis", False))
        before = xi.clone()     # (device copy; fetched only if somebody looks at cost_traj)
        out = self.engine.step(cfg, xi, start, end, rows, update=update_mode, want_grad=True, debug=want_dbg,
                               want_row_obs=True)
        B, n = xi.shape[0], xi.shape[1]
        host_xi, register = self._fresh_pinned_result(xi)
        host_info = self._stage.get(("host_info", B))
        if host_info is None:
            host_info = torch.empty((B, out["info"].shape[1]), dtype=torch.float64).pin_memory()
            self._stage[("host_info", B)] = host_info
        host_info.copy_(out["info"], non_blocking=True)
        grad, row_obs = out["grad"], out["row_obs"]
        if not batched:
            # one trajectory: its info dict is built right away (the reference's shape), so the gradient, the per-row
            # obstacle costs and the previous state ride back behind the same synchronisation
            small = self._stage.get(("host_small", n))
            if small is None:
                small = tuple(torch.empty(shp, dtype=torch.float64).pin_memory() for shp in ((1, n, 9), (1, n), (1, n, 9)))
                self._stage[("host_small", n)] = small
            for dst, src in zip(small, (grad, row_obs, before)):
                dst.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        new_xi, info_rows = host_xi.numpy(), host_info.numpy().copy()
        if batched:
            register(new_xi)    # Optimizer.optimize hands this very array back as traj.data
        else:
            grad, row_obs, before = (t.numpy().copy() for t in small)
        st_h, en_h = self._stage["start"][0].numpy().copy(), self._stage["end"][0].numpy().copy()
        self._last_start, self._last_end = st_h, en_h
        infos = BatchInfos(self, cfg, info_rows, n, grad, row_obs, before, st_h, en_h)
        if want_dbg:
            self._fill_collision_pts(infos, out)
        return (infos if batched else [infos[0]]), new_xi, batched

    def _fill_collision_pts(self, infos, out):
        """info['collision_pts'] [n,10,p,12] (omg/cost.py:355-358): xyz, potential, potential gradient."""
        pts, pot = out["points"], out["potentials"]
        B, n, m, p = pot.shape
        poses, eps, pad, clr, dis = (torch.from_numpy(a).to(pts.device) for a in self.object_params())
        _, grads, _ = self.sdf_loss(poses, self.env.sdf_torch, self.env.sdf_limits.to(pts.device).float().contiguous(),
                                    pts.reshape(-1, 3).contiguous(), eps, pad, clr, dis)
        grads = grads.reshape(B, n, m, p, 3).cpu().numpy()
        for b, info in enumerate(infos):
            vis = np.zeros([n, m, p, 12])
            vis[..., :3] = pts[b].cpu().numpy(); vis[..., 6] = pot[b].cpu().numpy(); vis[..., 9:] = grads[b]
            info["collision_pts"] = vis

    def compute_total_loss(self, traj):
        """(cost, grad, info) like omg/cost.py:451-532 (no update)."""
        infos, _, batched = self.evaluate(traj, update_mode=0)
        if batched:
            return (np.array(infos.cost), infos.gradient(), infos)
        return infos[0]["cost"], infos[0]["gradient"], infos[0]

    def compute_obstacle_cost_layer(self, ws_positions, vis_pts=None, special_check_id=0,
                                    uncheck_finger_collision=-1, grad_free=True):
        """omg/cost.py:288-360: ws_positions [n,m,p,3] torch CUDA fp32 -> potentials, grads, collides."""
        n, m, p, _ = ws_positions.shape
        dev = ws_positions.device
        poses, eps, pad, clr, dis = (torch.from_numpy(a).to(dev) for a in self.object_params())
        pot, grad, col = self.sdf_loss(poses, self.env.sdf_torch, self.env.sdf_limits.float().contiguous(),
                                       ws_positions.reshape(-1, 3).contiguous().float(), eps, pad, clr, dis)
        pot, grad, col = pot.reshape(n, m, p), grad.reshape(n, m, p, 3), col.reshape(n, m, p)
        if uncheck_finger_collision == -1:
            pot[:, -2:] *= 0.1; grad[:, -2:] *= 0.1; col[:, -2:] = 0
        if vis_pts is not None:
            vis_pts[:, :m, :, :3] = ws_positions.detach().cpu().numpy()
            vis_pts[:, :m, :, 6] = pot.detach().cpu().numpy()
            vis_pts[:, :m, :, 9:] = grad.detach().cpu().numpy()
        return pot, grad, col

    def goal_costs(self, data, first, goals, uncheck_finger_collision=0):
        """The device half of Learner.cost_vector (omg/online_learner.py:123-150) in one fused launch:
        data [B,n,9] (traj.data), `first` the waypoint the straight lines start from, goals [B,G,9] or [1,G,9] /
        [G,9] (shared) -> numpy fp32 [B,G] = sum of potential x workspace speed along each line."""
        self.sync()
        dev = self.engine.device
        xi = torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float64))).to(dev)
        g = np.asarray(goals, dtype=np.float64)
        if g.ndim == 3 and g.shape[0] == 1 and xi.shape[0] > 1:
            g = g[0]
        g = torch.from_numpy(np.ascontiguousarray(g)).to(dev)
        return self.engine.goal_costs(xi, int(first), g, float(self.cfg.time_interval),
                                      uncheck_finger_collision).cpu().numpy()

    def batch_obstacle_cost(self, joints, arc_length=-1, only_collide=False, special_check_id=0,
                            uncheck_finger_collision=-1, start=None, end=None):
        """omg/cost.py:192-286.  Returns (potentials [M,10,p], grad [M,10,p,3], vis_pts, collide) as CUDA
        tensors (vis_pts: numpy [M,10,p,12] with xyz unset unless cfg.vis)."""
        self.sync()
        dev = self.engine.device
        q = torch.from_numpy(np.ascontiguousarray(np.asarray(joints, dtype=np.float64).reshape(-1, 9))).to(dev)
        st = None
        if arc_length > 0:
            st = torch.from_numpy(np.ascontiguousarray(np.asarray(start, dtype=np.float64).reshape(9))).to(dev)
        pot, grad, col = self.engine.batch_obstacle_cost(q, arc_length, st, float(self.cfg.time_interval),
                                                         uncheck_finger_collision, want_grad=True)
        if only_collide:   # cost.py:279-284
            thr = 0.5 * (self.cfg.epsilon - self.cfg.clearance) ** 2 / self.cfg.epsilon
            pot = pot * (pot > thr).any()
        vis_pts = np.zeros([pot.shape[0], pot.shape[1], pot.shape[2], 12])
        return pot, grad, vis_pts, col
