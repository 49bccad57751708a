"""CPU: pin oracle/sdf_loss_ref.c (restatement of layers/sdf_matching_loss_kernel.cu) on fields whose
trilinear interpolant is known in closed form, and on the operator's documented quirks
(SURVEY.md Appendix A items 7, 8, 17, 20)."""
import numpy as np

from oracle import sdf_loss_ref as op


def linear_scene(coef, const, dims=(12, 10, 14), delta=0.05, origin=(-0.3, -0.25, -0.35)):
    """grid value at voxel (i,j,k) = coef . centre + const, centre = origin + (idx+0.5)*delta."""
    X, Y, Z = dims
    c = [origin[a] + (np.arange(d) + 0.5) * delta for a, d in enumerate(dims)]
    g = (coef[0] * c[0][:, None, None] + coef[1] * c[1][None, :, None] + coef[2] * c[2][None, None, :] + const)
    lim = np.array([[origin[0], origin[1], origin[2], origin[0] + X * delta, origin[1] + Y * delta,
                     origin[2] + Z * delta, X, Y, Z, delta]], np.float32)
    return g[None].astype(np.float32), lim


def run(grid, lim, pts, pose=None, eps=0.2, pad=1.0, clr=0.01, dis=0.0):
    pose = np.eye(4, dtype=np.float32)[None] if pose is None else pose
    o = grid.shape[0]
    f = lambda v: np.full(o, v, np.float32)
    return op.sdf_loss_forward(pose, grid, lim, pts, f(eps), f(pad), f(clr), f(dis), return_pin=True)


def test_linear_field_value_and_gradient():
    coef, const = np.array([0.3, -0.2, 0.5]), 0.02
    grid, lim = linear_scene(coef, const)
    rng = np.random.RandomState(0)
    pts = rng.uniform([-0.15, -0.1, -0.2], [0.15, 0.1, 0.2], (200, 3)).astype(np.float32)
    pot, grad, col, pin = run(grid, lim, pts)
    v = pts.astype(np.float64) @ coef + const
    assert pin == 200
    exp_pot = np.where(v <= 0, -v + 0.1, np.where(v <= 0.2, (v - 0.2) ** 2 / 0.4, 0.0))
    np.testing.assert_allclose(pot, exp_pot, atol=2e-6)
    scale = np.where(v <= 0, -1.0, np.where(v <= 0.2, (v - 0.2) / 0.2, 0.0))
    np.testing.assert_allclose(grad, scale[:, None] * coef[None], atol=3e-5)
    np.testing.assert_array_equal(col, (v < 0.01).astype(np.float32))


def test_out_of_bounds_is_far_and_never_collides():
    grid, lim = linear_scene(np.zeros(3), -1.0)  # deep inside everywhere in the grid
    pts = np.array([[5, 0, 0], [0, -5, 0], [0.299, 0, 0], [-0.33, 0, 0]], np.float32)
    pot, grad, col, pin = run(grid, lim, pts)
    # third: inside the box but its 8-tap cell sticks out (x1 == dim); fourth: (int)(g-0.5) == -1
    assert pin == 0
    assert not pot.any() and not grad.any() and not col.any()


def test_negative_half_cell_extrapolates():
    """(int) truncation: grid coordinate in (-0.5, 0.5) maps to cell 0 with a negative weight (A-7)."""
    coef, const = np.array([1.0, 0.0, 0.0]), 0.0
    grid, lim = linear_scene(coef, const)
    x = -0.3 + 0.2 * 0.05  # grid coordinate 0.2 -> x0 = (int)(-0.3) = 0, fx = -0.3
    pts = np.array([[x, 0.0, 0.0]], np.float32)
    pot, grad, col, pin = run(grid, lim, pts)
    assert pin == 1
    v = float(np.float32(x))  # linear field: extrapolation is exact
    np.testing.assert_allclose(pot[0], -v + 0.1, atol=1e-6)
    # gradient: the -x sample is OOB -> 1.0 mixes in (A-7): 0.5*(f(+1) - 1.0)/delta
    fp = v + 0.05
    np.testing.assert_allclose(grad[0, 0], -(0.5 * (fp - 1.0) / 0.05), rtol=1e-5)


def test_pose_rotation_sum_over_objects_disable_and_padding():
    coef, const = np.array([0.0, 0.0, 1.0]), 0.05
    g1, l1 = linear_scene(coef, const)
    grid = np.concatenate([g1, g1, g1]); lim = np.concatenate([l1, l1, l1])
    a = 0.7
    rot = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    pose = np.tile(np.eye(4, dtype=np.float32), (3, 1, 1))
    pose[1, :3, :3] = rot.astype(np.float32); pose[1, :3, 3] = [0.01, -0.02, 0.03]
    pts = np.array([[0.02, 0.03, 0.04], [-0.05, 0.01, -0.02]], np.float32)
    eps = np.array([0.2, 0.2, 0.2], np.float32); pad = np.array([1, 0.5, 1], np.float32)
    clr = np.array([0.01, 0.2, 0.01], np.float32); dis = np.array([0, 0, 1], np.float32)
    pot, grad, col = op.sdf_loss_forward(pose, grid, lim, pts, eps, pad, clr, dis)
    exp_pot = np.zeros(2); exp_grad = np.zeros((2, 3)); exp_col = np.zeros(2)
    for o in range(2):  # object 2 disabled
        R, t = pose[o, :3, :3].astype(np.float64), pose[o, :3, 3].astype(np.float64)
        q = pts.astype(np.float64) @ R.T + t
        v = q[:, 2] + const
        assert ((v > 0) & (v <= 0.2)).all()
        exp_pot += (v - 0.2) ** 2 / 0.4 * pad[o]
        exp_grad += ((v - 0.2) / 0.2 * pad[o])[:, None] * (R.T @ coef)[None]
        exp_col += v < clr[o]
    np.testing.assert_allclose(pot, exp_pot, atol=2e-6)
    np.testing.assert_allclose(grad, exp_grad, atol=3e-5)
    np.testing.assert_array_equal(col, exp_col)


def test_trilinear_matches_manual_on_random_grid():
    rng = np.random.RandomState(3)
    dims, delta, origin = (9, 8, 7), 0.1, np.array([-0.45, -0.4, -0.35])
    g = rng.uniform(0.3, 0.6, dims).astype(np.float32)  # > eps: potential 0, but collide uses value
    lim = np.array([[*origin, *(origin + np.array(dims) * delta), *dims, delta]], np.float32)
    pts = rng.uniform(-0.2, 0.2, (50, 3)).astype(np.float32)
    thr = 0.45
    pot, grad, col, pin = run(g[None], lim, pts, clr=thr)
    gc = (pts.astype(np.float64) - origin) / delta - 0.5
    i0 = np.floor(gc).astype(int); f = gc - i0
    val = np.zeros(50)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (f[:, 2] if dz else 1 - f[:, 2])
                val += w * g[i0[:, 0] + dx, i0[:, 1] + dy, i0[:, 2] + dz]
    assert pin == 50 and not pot.any()
    sure = np.abs(val - thr) > 1e-5
    np.testing.assert_array_equal(col[sure], (val < thr)[sure].astype(np.float32))


# ---- pin against the reference's own interpolation helpers (layers/sdf_matching_loss_kernel.cu:15-86) ------------
def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_interp_equals_reference_fixture():
    """oracle/sdf_loss_ref.c's value_interp / grad_interp == outputs of the reference's getValueInterpolated /
    getGradientInterpolated (compiled from the reference source; tools/make_golden_sdf_interp.py), bit for bit."""
    import os
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "sdf_interp.npz"))
    val, grad = op.interp(fx["pgrid"], fx["grid"], float(fx["delta"]))
    assert (fx["value"] == 1.0).sum() > 1000 and (fx["value"] != 1.0).sum() > 5000
    np.testing.assert_array_equal(_bits(val), _bits(fx["value"]))
    np.testing.assert_array_equal(_bits(grad), _bits(fx["grad"]))


def test_interp_equals_reference_helpers():
    """Live against oracle/_ref/libsdf_ref.so (the reference source compiled by oracle/sdf_ref/Makefile): 2 x 10^5
    random grid coordinates on two grids, incl. the (-0.5, 0.5) truncation band, border voxels and OOB taps."""
    import pytest
    from oracle import sdf_ref_lib
    if not sdf_ref_lib.have_ref():
        pytest.skip("oracle/_ref/libsdf_ref.so not built (needs /root/reference)")
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from make_golden_sdf_interp import coordinates
    for seed, dims in ((1, (13, 9, 21)), (2, (40, 33, 17))):
        rng = np.random.RandomState(seed)
        grid = rng.normal(0.1, 0.3, dims).astype(np.float32)
        pg = coordinates(dims, 100000, rng)
        delta = np.float32(rng.uniform(0.002, 0.02))
        v_ref, g_ref = sdf_ref_lib.interp(pg, grid, delta)
        v, g = op.interp(pg, grid, delta)
        np.testing.assert_array_equal(_bits(v), _bits(v_ref))
        np.testing.assert_array_equal(_bits(g), _bits(g_ref))
        band = ((pg > -0.5) & (pg < 0.5)).any(1)
        assert band.sum() > 1000 and (v_ref[band] != 1.0).sum() > 100   # the truncation band is exercised in bounds
