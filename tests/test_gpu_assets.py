"""GPU: trajectory initialisation (omgb_traj_interpolate) and the SDF asset path (omgb_sdf_pack, omgb_point_sdf)
through the host mirrors, against the fixtures produced by the reference's own code and against the oracle at
BASELINE sizes.  Byte/voxel work is bit-exact; the spline is fp64 arithmetic in a different association order than
scipy's LAPACK solve, asserted to 1e-12 rad."""
import os

import numpy as np
import pytest
import torch

from omg_planner_b200 import core as C
from omg_planner_b200.config import ChompConfig
from omg_planner_b200.sdf_tools import SignedDensityField
from oracle import sdf_asset_ref as A
from oracle import traj_ref as T

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def test_interpolate_matches_reference_fixtures():
    g = np.load(os.path.join(GOLD, "assets_traj.npz"))
    worst = 0.0
    for k in [k for k in g.files if k.startswith("out_")]:
        _, K, n, mode = k.split("_")
        out = C.interpolate_waypoints_device(_dev(g["wp_" + k[4:]][None]), int(n[1:]), mode)[0].cpu().numpy()
        worst = max(worst, np.abs(out - g[k]).max())
        np.testing.assert_allclose(out, g[k], rtol=0, atol=TOL, err_msg=k)
    print("worst |xi - reference| over 32 cases: %.2e" % worst)


def test_trajectory_mirror_fixed_and_dynamic_timesteps():
    g = np.load(os.path.join(GOLD, "assets_traj.npz"))
    cfg = ChompConfig(goal_set_proj=False)
    t = C.Trajectory(30, cfg=cfg, start=g["starts"], end=g["ends"])         # batched extension
    np.testing.assert_allclose(t.data, g["fixed"], rtol=0, atol=TOL)
    for b in range(g["starts"].shape[0]):                                   # reference shape: one trajectory
        cfg = ChompConfig(goal_set_proj=False, dynamic_timestep=True)
        tb = C.Trajectory(30, cfg=cfg, start=g["starts"][b], end=g["ends"][b])
        n = int(g["dynamic_n"][b])
        assert tb.data.shape == (n, 9) and cfg.timesteps == n and cfg.Ainv.shape == (n, n)
        np.testing.assert_allclose(tb.data, g["dynamic"][b, :n], rtol=0, atol=TOL)
    with pytest.raises(RuntimeError):                                       # mixed waypoint counts in one batch
        C.Trajectory(30, cfg=ChompConfig(goal_set_proj=False, dynamic_timestep=True), start=g["starts"], end=g["ends"])


def test_interpolate_full_batch_vs_oracle_and_properties():
    rng = np.random.RandomState(2)
    for K, n in ((2, 30), (2, 60), (5, 50)):
        wp = rng.uniform(-2.8, 2.8, (1024, K, 9))
        out = C.interpolate_waypoints_device(_dev(wp), n, "cubic").cpu().numpy()
        for b in (0, 17, 1023):
            np.testing.assert_allclose(out[b], T.interpolate_waypoints(wp[b], n), rtol=0, atol=TOL)
        lin = C.interpolate_waypoints_device(_dev(wp), n, "linear").cpu().numpy()
        if K == 2:
            # size-independent properties: symmetric blend (reversing the knots reverses the samples), bounded by
            # the end points, linear mode is the chord
            rev = C.interpolate_waypoints_device(_dev(wp[:, ::-1]), n, "cubic").cpu().numpy()
            np.testing.assert_allclose(rev, out[:, ::-1], rtol=0, atol=1e-13)
            lo, hi = wp.min(1)[:, None], wp.max(1)[:, None]
            assert (out >= lo - 1e-12).all() and (out <= hi + 1e-12).all()
            t = np.linspace(0, 1, n + 2)[1:-1][None, :, None]
            np.testing.assert_allclose(lin, wp[:, :1] * (1 - t) + wp[:, 1:] * t, rtol=0, atol=1e-13)
    assert C.interpolate_waypoints_device(_dev(np.zeros((0, 2, 9))), 30).shape == (0, 30, 9)


def _write_pth(path, stored, min_coords, delta):
    shape = (stored.shape[1], stored.shape[0], stored.shape[2])
    torch.save({"min_coords": torch.from_numpy(np.array(min_coords)),
                "max_coords": torch.from_numpy(np.array(min_coords) + delta * np.array(shape)),
                "delta": float(delta), "sdf_torch": torch.from_numpy(stored)[None, None]}, path)


def test_from_pth_resize_combine_sdfs_bit_exact(tmp_path):
    g = np.load(os.path.join(GOLD, "assets_sdf.npz"))
    objs = []
    for i in range(int(g["num"])):
        p = str(tmp_path / ("obj%d.pth" % i))
        _write_pth(p, g["stored%d" % i], g["mins"][i], float(g["deltas"][i]))
        f = SignedDensityField.from_pth(p)
        f.resize(float(g["ratios"][i]))
        np.testing.assert_array_equal(f.data_torch.cpu().numpy(), g["data_torch%d" % i])
        np.testing.assert_array_equal(f.data, g["data_torch%d" % i])
        objs.append(type("Obj", (), {"sdf": f, "name": "obj%d" % i})())
    env = type("Env", (), {"objects": objs})()
    grids, limits = C.combine_sdfs(env)
    assert grids.is_cuda and grids.dtype == torch.float32 and limits.is_cuda
    np.testing.assert_array_equal(grids.cpu().numpy(), g["combined"])
    np.testing.assert_array_equal(limits.cpu().numpy(), g["limits"])


def test_sdf_pack_config2_size_vs_oracle():
    """10 objects, mixed shapes up to 128^3 (odd z extents exercise the scalar store path), fp32 and fp64 sources,
    both layouts."""
    rng = np.random.RandomState(4)
    for zmax in (128, 127):
        fields, refs = [], []
        for i in range(10):
            shp = (128, 128, zmax) if i == 0 else tuple(int(v) for v in rng.randint(40, 129, 3))
            shp = (shp[0], shp[1], min(shp[2], zmax))
            data = rng.uniform(-0.1, 0.4, shp).astype(np.float32 if i % 2 == 0 else np.float64)
            origin = rng.uniform(-0.3, -0.1, 3)
            ref = A.FieldRef(data.copy(), origin.copy(), 0.004)
            if i % 3 == 0:   # stored the .pth way
                raw = torch.from_numpy(np.ascontiguousarray(data.transpose(1, 0, 2))).cuda()
                f = SignedDensityField(shp, origin.copy(), 0.004, _raw=raw, _layout=1)
            else:
                f = SignedDensityField(data, origin.copy(), 0.004)
            if i % 4 == 1:
                f.resize(0.9); ref.resize(0.9)
            fields.append(f); refs.append(ref)
        want_g, want_l = A.combine_sdfs(refs)
        mx = np.array([f.shape for f in fields]).max(0)
        got = C.pack_sdf_grids(fields, mx).cpu().numpy()
        np.testing.assert_array_equal(got, want_g)
        np.testing.assert_array_equal(C.sdf_limits_for(fields, mx), want_l)


def test_point_sdf_bit_exact_vs_reference_fixture_and_oracle():
    g = np.load(os.path.join(GOLD, "assets_sdf.npz"))
    for tag in ("cloud", "empty"):
        f, d64 = C.compute_sdf_from_points(g[tag + "_points"], keep_fp64=True)
        np.testing.assert_array_equal(d64.cpu().numpy(), g[tag + "_dists"])
        np.testing.assert_array_equal(np.asarray(f.min_coords), g[tag + "_origin"])
        env = type("Env", (), {"objects": [type("Obj", (), {"sdf": f})()]})()
        grids, limits = C.combine_sdfs(env)
        np.testing.assert_array_equal(grids.cpu().numpy(), g[tag + "_sdf_torch"])
        np.testing.assert_array_equal(limits.cpu().numpy(), g[tag + "_limits"])
    # a table-top sized cloud: 2500 points, ~50^3 voxels, several shared-memory tiles with a ragged last tile
    rng = np.random.RandomState(12)
    pts = rng.uniform([0.3, -0.3, 0.0], [0.8, 0.3, 0.4], (2500, 3))
    f, d64 = C.compute_sdf_from_points(pts, keep_fp64=True)
    want, origin, _ = A.point_sdf(pts)
    assert d64.shape == want.shape
    np.testing.assert_array_equal(d64.cpu().numpy(), want)
    # property: every cloud point's own voxel neighbourhood is within half a voxel diagonal of it
    idx = np.floor((pts - origin) / 0.02 + 0.5).astype(int)
    near = d64.cpu().numpy()[idx[:, 0], idx[:, 1], idx[:, 2]]
    assert (near <= 0.02 * np.sqrt(3) / 2 + 1e-12).all()
