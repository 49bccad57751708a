#!/usr/bin/env python
"""Generate tests/golden/assets_traj.npz and assets_sdf.npz by running the UNMODIFIED reference code
(/root/reference under the stubs of tools/ref_harness.py):

  assets_traj  omg.util.interpolate_waypoints (cubic / linear, 2..6 knots) and omg.core.Trajectory (fixed and dynamic
               timesteps)
  assets_sdf   omg.sdf_tools.SignedDensityField.from_pth + .resize on .pth files written the way
               real_world/convert_sdf.py:43-77 writes them, omg.core.Env.combine_sdfs, and
               omg.core.PointEnv.compute_sdf_from_points (scipy cKDTree)

BUILD-CONTAINER ONLY (needs /root/reference).  The committed .npz files are what travels.
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as H  # noqa: E402


def load_core():
    ns = H.load_reference()
    sys.modules.setdefault("ycb_render.ycb_renderer", H._Anything("ycb_render.ycb_renderer"))
    ns.core = importlib.import_module("omg.core")
    ns.sdf_tools = importlib.import_module("omg.sdf_tools")
    return ns


def make_traj(ns, out_dir):
    rng = np.random.RandomState(5)
    cases = {}
    for K in (2, 3, 4, 6):
        for n in (1, 7, 30, 60):
            wp = rng.uniform(-2.5, 2.5, (K, 9))
            for mode in ("cubic", "linear"):
                cases["wp_K%d_n%d_%s" % (K, n, mode)] = wp
                cases["out_K%d_n%d_%s" % (K, n, mode)] = ns.util.interpolate_waypoints(wp, n, 9, mode=mode)
    # Trajectory objects: fixed timesteps, then dynamic timesteps (omg/core.py:59-78)
    cfg = ns.cfg
    starts = rng.uniform(-2.0, 2.0, (5, 9)); ends = rng.uniform(-2.0, 2.0, (5, 9))
    ends[1] = starts[1] + 0.4           # -> 24 waypoints
    ends[2] = starts[2] - 0.27          # -> 16 waypoints
    ends[3] = starts[3] + 0.01          # -> traj_min_step
    ends[4] = starts[4] + 3.0           # -> traj_max_step
    fixed, dyn, dyn_n = [], [], []
    for b in range(5):
        cfg.dynamic_timestep = False
        cfg.timesteps = 30
        cfg.get_global_param(30)
        t = ns.core.Trajectory(30)
        t.start, t.end = starts[b].copy(), ends[b].copy()
        t.interpolate_waypoints(mode=cfg.traj_interpolate)
        fixed.append(t.data.copy())
        cfg.dynamic_timestep = True
        t.interpolate_waypoints(mode=cfg.traj_interpolate)
        dyn.append(t.data.copy()); dyn_n.append(cfg.timesteps)
        cfg.dynamic_timestep = False
        cfg.timesteps = 30
        cfg.get_global_param(30)
    pad = np.zeros((5, 50, 9))
    for b in range(5):
        pad[b, :dyn_n[b]] = dyn[b]
    path = os.path.join(out_dir, "assets_traj.npz")
    np.savez_compressed(path, starts=starts, ends=ends, fixed=np.stack(fixed), dynamic=pad, dynamic_n=np.array(dyn_n),
                        traj_delta=cfg.traj_delta, traj_min_step=cfg.traj_min_step, traj_max_step=cfg.traj_max_step,
                        **cases)
    print("->", path, os.path.getsize(path) // 1024, "KiB; dynamic timesteps", dyn_n)


def make_sdf(ns, out_dir):
    rng = np.random.RandomState(9)
    shapes = [(12, 10, 14), (16, 16, 16), (9, 20, 11), (5, 6, 7)]
    ratios = [1.0, 1.0, 0.8, 1.25]      # cfg.target_size values passed to resize (omg/core.py:109)
    tmp = tempfile.mkdtemp()
    objects, stored, mins, deltas = [], [], [], []
    for i, shp in enumerate(shapes):
        sdf = rng.uniform(-0.05, 0.3, shp)                                # float64 grid read from the .sdf text file
        min_coords = rng.uniform(-0.2, -0.05, 3)
        delta = float(rng.uniform(0.004, 0.012))
        # real_world/convert_sdf.py:43-77, verbatim recipe
        sdf_torch = torch.from_numpy(sdf).float().permute(1, 0, 2).unsqueeze(0).unsqueeze(1)
        max_coords = min_coords + delta * np.array(sdf.shape)
        path = os.path.join(tmp, "obj%d.pth" % i)
        torch.save({"min_coords": torch.from_numpy(min_coords), "max_coords": torch.from_numpy(max_coords),
                    "delta": delta, "sdf_torch": sdf_torch}, path)
        stored.append(sdf_torch[0, 0].numpy().copy()); mins.append(min_coords.copy()); deltas.append(delta)
        o = types.SimpleNamespace(name="obj%d" % i)
        _load = torch.load
        torch.load = lambda f, *a, **k: _load(f, *a, weights_only=False, **k)   # (torch >= 2.6 default flipped)
        try:
            o.sdf = ns.sdf_tools.SignedDensityField.from_pth(path)
        finally:
            torch.load = _load
        o.sdf.resize(ratios[i])                                            # omg/core.py:109
        objects.append(o)
    env = types.SimpleNamespace(objects=objects)
    ns.cfg.report_time = False
    ns.core.Env.combine_sdfs(env)
    out = dict(num=len(shapes), ratios=np.array(ratios), deltas=np.array(deltas), mins=np.stack(mins),
               combined=env.sdf_torch.numpy(), limits=env.sdf_limits.numpy())
    for i in range(len(shapes)):
        out["stored%d" % i] = stored[i]
        out["data_torch%d" % i] = objects[i].sdf.data_torch.numpy()
    # PointEnv.compute_sdf_from_points, unmodified; only add_object (file loading) is replaced
    for tag, pts in (("cloud", rng.uniform([0.3, -0.1, 0.0], [0.42, 0.05, 0.1], (400, 3))), ("empty", np.zeros((0, 3)))):
        pe = ns.core.PointEnv.__new__(ns.core.PointEnv)
        pe.objects, pe.grid_resolution = [], 0.02
        pe.add_object = lambda *a, **k: pe.objects.append(types.SimpleNamespace(name="perception/env_points"))
        import builtins
        _print = builtins.print
        builtins.print = lambda *a, **k: None
        try:
            pe.compute_sdf_from_points(pts)
        finally:
            builtins.print = _print
        f = pe.objects[0].sdf
        out["%s_points" % tag] = pts
        out["%s_dists" % tag] = f.data
        out["%s_origin" % tag] = np.array(f.min_coords)
        out["%s_sdf_torch" % tag] = pe.sdf_torch.numpy()
        out["%s_limits" % tag] = pe.sdf_limits.numpy()
    path = os.path.join(out_dir, "assets_sdf.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) // 1024, "KiB; combined", out["combined"].shape, "cloud grid",
          out["cloud_dists"].shape, "empty grid", out["empty_dists"].shape)


if __name__ == "__main__":
    ns = load_core()
    out_dir = os.path.join(ROOT, "tests", "golden")
    make_traj(ns, out_dir)
    make_sdf(ns, out_dir)
