"""The data carriers either side of the CHOMP hot path, device-backed:

  Trajectory                 omg/core.py:23-78   -- initialisation by omgb_traj_interpolate (one launch per batch)
  combine_sdfs / pack_sdf_grids  omg/core.py:366-411 -- omgb_sdf_pack (one launch per scene)
  compute_sdf_from_points    omg/core.py:426-457 -- omgb_point_sdf (nearest-point distance field of a point cloud)

Everything numeric runs in libomgb200.so; there is no CPU fallback."""
import ctypes

import numpy as np
import torch

from . import _lib
from .sdf_tools import SignedDensityField

_vp = ctypes.c_void_p

START_CONF = np.array([0.0, -1.285, 0, -2.356, 0.0, 1.571, 0.785, 0.04, 0.04])   # omg/core.py:38
END_CONF = np.array([-0.99, -1.74, -0.61, -3.04, 0.88, 1.21, -1.12, 0.04, 0.04])  # omg/core.py:39


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("omg_planner_b200.core needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


# ---- trajectory initialisation ---------------------------------------------------------------------------
def interpolate_waypoints_device(waypoints, n, mode="cubic"):
    """omg/util.py:238-258 for a batch: waypoints torch CUDA fp64 [B,K,9] -> [B,n,9] (one launch)."""
    if not (waypoints.is_cuda and waypoints.dtype == torch.float64 and waypoints.is_contiguous()
            and waypoints.dim() == 3 and waypoints.shape[2] == 9):
        raise RuntimeError("waypoints must be a contiguous fp64 CUDA tensor [B,K,9]")
    if mode not in ("cubic", "linear"):
        raise RuntimeError("mode must be 'cubic' or 'linear'")   # (the reference's "quintic" branch is a no-op)
    B, K = waypoints.shape[0], waypoints.shape[1]
    xi = torch.empty((B, int(n), 9), dtype=torch.float64, device=waypoints.device)
    with torch.cuda.device(waypoints.device):
        _lib.check(_lib.lib().omgb_traj_interpolate(_vp(waypoints.data_ptr()), B, K, int(n),
                                                    1 if mode == "cubic" else 0, _vp(xi.data_ptr()), _stream()),
                   "omgb_traj_interpolate")
    return xi


def dynamic_timesteps(start, end, cfg):
    """omg/core.py:64-72: waypoint count from the joint-space distance, per trajectory (vectorised)."""
    d = np.linalg.norm(np.asarray(start, dtype=np.float64) - np.asarray(end, dtype=np.float64), axis=-1)
    steps = (d / cfg.traj_delta).astype(int)
    return np.minimum(np.maximum(steps, cfg.traj_min_step), cfg.traj_max_step)


class Trajectory(object):
    """omg/core.py:23-78.  `.data` is [n,9] like the reference, or [B,n,9] when start/end are [B,9] (batched
    extension); `.data` is a numpy array (the plugin surface's type), produced on the device."""

    def __init__(self, timesteps=100, dof=9, cfg=None, start=None, end=None):
        if cfg is None:
            raise RuntimeError("Trajectory needs the scene's cfg (the reference reads the global config.cfg)")
        self.cfg = cfg
        self.timesteps = cfg.timesteps
        self.dof = dof
        self.goal_set = []
        self.goal_quality = []
        self.goal_idx = 0
        self.start = START_CONF.copy() if start is None else np.array(start, dtype=np.float64)
        self.end = END_CONF.copy() if end is None else np.array(end, dtype=np.float64)
        self.data = np.zeros(self.start.shape[:-1] + (self.timesteps, dof))
        self.interpolate_waypoints(mode=getattr(cfg, "traj_interpolate", "cubic"))

    def update(self, grad):
        """omg/core.py:43-51."""
        if self.cfg.consider_finger:
            self.data += grad
        else:
            self.data[..., :-2] += grad[..., :-2]
        self.data[..., -2:] = np.minimum(np.maximum(self.data[..., -2:], 0), 0.04)

    def set(self, new_traj):
        self.data = new_traj

    def interpolate_waypoints(self, waypoints=None, mode="cubic"):
        """omg/core.py:59-78 (`waypoints` is ignored there too: the knots are always start and end)."""
        cfg = self.cfg
        timesteps = cfg.timesteps
        start, end = np.asarray(self.start, dtype=np.float64), np.asarray(self.end, dtype=np.float64)
        if getattr(cfg, "dynamic_timestep", False):
            steps = np.atleast_1d(dynamic_timesteps(start, end, cfg))
            if not (steps == steps[0]).all():
                raise RuntimeError("dynamic_timestep: trajectories of one batch must share a waypoint count; "
                                   "bucket them by core.dynamic_timesteps first")
            timesteps = int(steps[0])
            cfg.timesteps = timesteps
            cfg.get_global_param(timesteps)
            self.timesteps = timesteps
        batched = start.ndim == 2
        s2, e2 = np.atleast_2d(start), np.atleast_2d(end)
        e2 = np.broadcast_to(e2, s2.shape) if e2.shape != s2.shape else e2
        wp = torch.from_numpy(np.ascontiguousarray(np.stack([s2, e2], axis=1))).to(_device())
        xi = interpolate_waypoints_device(wp, timesteps, mode).cpu().numpy()
        self.data = xi if batched else xi[0]


# ---- SDF packing ---------------------------------------------------------------------------------------
def pack_sdf_grids(sdfs, max_shape, device=None):
    """The tensor half of Env.combine_sdfs (omg/core.py:372-379): every SignedDensityField's raw device grid ->
    its corner of a [O,X,Y,Z] fp32 tensor padded with 1.0, in ONE launch (permute / resize / fp64->fp32 fused)."""
    dev = _device(device)
    num = len(sdfs)
    X, Y, Z = (int(v) for v in max_shape)
    out = torch.empty((num, X, Y, Z), dtype=torch.float32, device=dev)
    if num == 0:
        return out
    table = (_lib.SdfSource * num)()
    for i, f in enumerate(sdfs):
        raw = f.raw
        if not (raw.is_cuda and raw.is_contiguous() and raw.device == dev):
            raise RuntimeError("SignedDensityField.raw must be a contiguous tensor on %s" % dev)
        table[i].data = raw.data_ptr()
        table[i].shape[0], table[i].shape[1], table[i].shape[2] = f.nx, f.ny, f.nz
        table[i].layout = f.layout
        table[i].dtype = 1 if raw.dtype == torch.float64 else 0
        table[i].scale = float(f.scale)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().omgb_sdf_pack(table, num, X, Y, Z, _vp(out.data_ptr()), _stream()), "omgb_sdf_pack")
    return out


def sdf_limits_for(sdfs, max_shape):
    """The [O,10] half of Env.combine_sdfs (omg/core.py:376-391), the reference's own expressions."""
    num = len(sdfs)
    limits = np.zeros((num, 10), dtype=np.float32)
    for i, f in enumerate(sdfs):
        size = f.shape
        xmins, ymins, zmins = f.min_coords
        xmaxs, ymaxs, zmaxs = f.max_coords
        limits[i, 0] = xmins
        limits[i, 1] = ymins
        limits[i, 2] = zmins
        limits[i, 3] = xmins + (xmaxs - xmins) * max_shape[0] / size[0]
        limits[i, 4] = ymins + (ymaxs - ymins) * max_shape[1] / size[1]
        limits[i, 5] = zmins + (zmaxs - zmins) * max_shape[2] / size[2]
        limits[i, 6] = max_shape[0]
        limits[i, 7] = max_shape[1]
        limits[i, 8] = max_shape[2]
        limits[i, 9] = f.delta
    return limits


def combine_sdfs(env, device=None):
    """Env.combine_sdfs (omg/core.py:366-411): sets env.sdf_torch [O,X,Y,Z] fp32 CUDA and env.sdf_limits [O,10]
    fp32 CUDA from env.objects[i].sdf."""
    sdfs = [obj.sdf for obj in env.objects]
    max_shape = np.array([f.shape for f in sdfs]).max(axis=0)
    env.sdf_torch = pack_sdf_grids(sdfs, max_shape, device)
    env.sdf_limits = torch.from_numpy(sdf_limits_for(sdfs, max_shape)).to(env.sdf_torch.device)
    return env.sdf_torch, env.sdf_limits


# ---- point-cloud SDF (PointEnv) --------------------------------------------------------------------------
def compute_sdf_from_points(points, grid_resolution=0.02, margin=0.24, device=None, keep_fp64=False):
    """PointEnv.compute_sdf_from_points (omg/core.py:426-452): unsigned nearest-point distance on a regular grid
    around the cloud's bounding box.  points [N,3] in the robot base frame -> SignedDensityField (plus the fp64
    distances when keep_fp64).  The reference builds a cKDTree and queries every voxel; here every voxel scans the
    cloud (staged through shared memory), in fp64 with cKDTree's operation order -- bit-identical distances."""
    dev = _device(device)
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    if points.shape[0] == 0:
        points = np.ones((2, 3)) * 3                                          # core.py:434-435
    bounds = np.stack((points.min(0), points.max(0)), axis=1)                 # workspace_bounds [3,2]
    axes = [np.arange(bounds[k][0] - margin, bounds[k][1] + margin, grid_resolution) for k in range(3)]
    X, Y, Z = (len(a) for a in axes)
    d_pts = torch.from_numpy(np.ascontiguousarray(points)).to(dev)
    d_ax = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in axes]
    out64 = torch.empty((X, Y, Z), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().omgb_point_sdf(_vp(d_pts.data_ptr()), points.shape[0], _vp(d_ax[0].data_ptr()),
                                             _vp(d_ax[1].data_ptr()), _vp(d_ax[2].data_ptr()), X, Y, Z, None,
                                             _vp(out64.data_ptr()), _stream()), "omgb_point_sdf")
    # SignedDensityField(dists, workspace_bounds[:,0] - margin, grid_resolution) (core.py:452): the fp64 distances
    # stay on the device; the pack kernel rounds them to fp32 exactly like `data.astype(np.float32)`
    field = SignedDensityField((X, Y, Z), bounds[:, 0] - margin, grid_resolution, _raw=out64, _layout=0, device=dev)
    field.workspace_bounds = bounds
    return (field, out64) if keep_fp64 else field
