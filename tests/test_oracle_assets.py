"""CPU: the oracle restatements of trajectory initialisation (oracle/traj_ref.py) and of the SDF asset path
(oracle/sdf_asset_ref.py) against fixtures produced by the reference's own code
(tools/make_golden_assets.py -> tests/golden/assets_*.npz)."""
import os

import numpy as np

from oracle import sdf_asset_ref as A
from oracle import traj_ref as T

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_interpolate_waypoints_matches_reference():
    g = np.load(os.path.join(GOLD, "assets_traj.npz"))
    keys = [k for k in g.files if k.startswith("out_")]
    assert len(keys) == 32
    for k in keys:
        _, K, n, mode = k.split("_")
        out = T.interpolate_waypoints(g["wp_" + k[4:]], int(n[1:]), mode=mode)
        np.testing.assert_allclose(out, g[k], rtol=0, atol=1e-13, err_msg=k)


def test_trajectory_init_fixed_and_dynamic():
    g = np.load(os.path.join(GOLD, "assets_traj.npz"))
    np.testing.assert_allclose(T.trajectory_init(g["starts"], g["ends"], 30), g["fixed"], rtol=0, atol=1e-13)
    steps = T.dynamic_timesteps(g["starts"], g["ends"], float(g["traj_delta"]), int(g["traj_min_step"]),
                                int(g["traj_max_step"]))
    np.testing.assert_array_equal(steps, g["dynamic_n"])
    assert len(set(steps.tolist())) >= 4
    for b, n in enumerate(steps):
        np.testing.assert_allclose(T.trajectory_init(g["starts"][b], g["ends"][b], int(n))[0], g["dynamic"][b, :n],
                                   rtol=0, atol=1e-13)


def _fields(g):
    fields = []
    for i in range(int(g["num"])):
        f = A.FieldRef.from_stored(g["stored%d" % i], g["mins"][i].copy(), float(g["deltas"][i]))
        f.resize(float(g["ratios"][i]))
        fields.append(f)
    return fields


def test_from_pth_resize_combine_sdfs_bit_exact():
    g = np.load(os.path.join(GOLD, "assets_sdf.npz"))
    fields = _fields(g)
    for i, f in enumerate(fields):
        np.testing.assert_array_equal(f.data32, g["data_torch%d" % i])
    grids, limits = A.combine_sdfs(fields)
    np.testing.assert_array_equal(grids, g["combined"])
    np.testing.assert_array_equal(limits, g["limits"])
    # the quirk: resize leaves max_coords alone, so the stretched limits are NOT min + delta * padded shape
    i = 2
    assert abs(limits[i, 3] - (limits[i, 0] + limits[i, 9] * limits[i, 6])) > 1e-3


def test_point_sdf_matches_ckdtree_bit_exact():
    g = np.load(os.path.join(GOLD, "assets_sdf.npz"))
    for tag in ("cloud", "empty"):
        d, origin, _ = A.point_sdf(g[tag + "_points"])
        np.testing.assert_array_equal(d, g[tag + "_dists"])
        np.testing.assert_array_equal(origin, g[tag + "_origin"])
        grids, limits = A.combine_sdfs([A.FieldRef(d, origin, 0.02)])
        np.testing.assert_array_equal(grids, g[tag + "_sdf_torch"])
        np.testing.assert_array_equal(limits, g[tag + "_limits"])
