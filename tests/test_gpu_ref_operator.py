"""GPU: the product operator (omgb_sdf_loss behind omg_cuda.sdf_loss_forward) against THE REFERENCE'S OWN DEVICE CODE:
layers/sdf_matching_loss_kernel.cu compiled from the reference source for sm_100a (oracle/sdf_ref/Makefile ->
oracle/_ref/libsdf_ref.so, which travels to the GPU box) and run here through its own host function
sdf_loss_cuda_forward (kernel.cu:204-262).  Bit-exact wherever the reference itself is deterministic (its
sum over objects is an atomicAdd in arbitrary order, kernel.cu:186-195: sums of <= 2 non-zero terms do not depend
on the order)."""
import os

import numpy as np
import pytest
import torch

import helpers as H
from omg_planner_b200 import scene as S
from omg_planner_b200.engine import sdf_loss_forward
from oracle import sdf_ref_lib

pytestmark = pytest.mark.gpu


def _need_ref():
    if not sdf_ref_lib.have_ref():
        pytest.skip("oracle/_ref/libsdf_ref.so not built (needs /root/reference in the build container)")


def _bits(t):
    return t.contiguous().view(torch.int32)


def _rand_pose(rng):
    """world -> object pose, fp32, general rotation (both branches of the matrix -> quaternion conversion)."""
    a = rng.normal(size=(3, 3))
    q, _ = np.linalg.qr(a)
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    m = np.eye(4)
    m[:3, :3] = q
    m[:3, 3] = rng.uniform(-0.4, 0.4, 3)
    return m.astype(np.float32)


def _scene_inputs(seed, num_objects, grid, n_points, big_eps=False):
    sc = S.make_scene(num_objects=num_objects, grid=grid, seed=seed)
    rng = np.random.RandomState(seed)
    O = len(sc["names"])
    pose = np.stack([_rand_pose(rng) for _ in range(O)])
    # points: around every object's origin in its own frame, mapped to the world (so that many samples are in bounds)
    pts = []
    for o in range(O):
        ext = (sc["sdf_limits"][o, 3:6] - sc["sdf_limits"][o, 0:3]) * 0.6
        local = rng.uniform(-ext, ext, (n_points // O, 3))
        R, t = pose[o, :3, :3].astype(np.float64), pose[o, :3, 3].astype(np.float64)
        pts.append((local - t) @ R)   # R^T (local - t)
    pts = np.concatenate(pts).astype(np.float32)
    eps = np.full(O, 0.6 if big_eps else 0.2, np.float32)
    eps[0] = 0.1
    pad = rng.uniform(0.5, 1.0, O).astype(np.float32)
    clr = np.full(O, 0.01, np.float32)
    dis = np.zeros(O, np.float32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return [dev(pose), dev(sc["sdf_grids"]), dev(sc["sdf_limits"]), dev(pts), dev(eps), dev(pad), dev(clr), dev(dis)]


def test_single_objects_bit_exact():
    """One object per call: the reference's atomic sum has a single term, everything is deterministic."""
    _need_ref()
    total = nonzero = 0
    for seed in range(4):
        args = _scene_inputs(seed, num_objects=5, grid=48, n_points=40000, big_eps=(seed % 2 == 1))
        O = args[0].shape[0]
        for o in range(O):
            one = [args[0][o:o + 1].contiguous(), args[1][o:o + 1].contiguous(), args[2][o:o + 1].contiguous(), args[3]] + \
                  [a[o:o + 1].contiguous() for a in args[4:]]
            ref = sdf_ref_lib.forward_device(*one)
            got = sdf_loss_forward(*one)
            for r, g, name in zip(ref, got, ("potentials", "potential_grads", "collides")):
                bad = (_bits(r) != _bits(g))
                assert not bad.any(), "%s: %d of %d differ (seed %d object %d)" % (name, int(bad.sum()), bad.numel(), seed, o)
            total += ref[0].numel()
            nonzero += int((ref[0] != 0).sum())
    assert nonzero > 0.05 * total   # the comparison is not vacuous


def test_scene_sum_over_objects():
    """All objects in one call.  Points to which at most two objects contribute are order-independent -> bit-exact;
    the rest differ by the reference's own run-to-run atomic order (<= a few ulp)."""
    _need_ref()
    for seed in (0, 1):
        args = _scene_inputs(10 + seed, num_objects=6, grid=40, n_points=60000, big_eps=True)
        ref = sdf_ref_lib.forward_device(*args)
        got = sdf_loss_forward(*args)
        # number of contributing objects per point, from single-object calls of the product operator
        O = args[0].shape[0]
        contrib = torch.zeros_like(ref[0])
        for o in range(O):
            one = [args[0][o:o + 1].contiguous(), args[1][o:o + 1].contiguous(), args[2][o:o + 1].contiguous(), args[3]] + \
                  [a[o:o + 1].contiguous() for a in args[4:]]
            contrib += (sdf_loss_forward(*one)[0] != 0).float()
        few = contrib <= 2
        assert int(few.sum()) > 1000 and int((~few).sum()) > 100
        assert torch.equal(_bits(ref[0])[few], _bits(got[0])[few])
        assert torch.equal(_bits(ref[1])[few], _bits(got[1])[few])
        assert torch.equal(ref[2], got[2])   # collide counts are small integers: exact in any order
        torch.testing.assert_close(got[0], ref[0], rtol=4e-7, atol=1e-7)
        torch.testing.assert_close(got[1], ref[1], rtol=1e-6, atol=2e-6)


def test_disabled_objects_and_oob():
    _need_ref()
    args = _scene_inputs(3, num_objects=4, grid=32, n_points=8000)
    args[7][1] = 1.0   # disables
    args[3][:100] += 50.0   # far outside every grid
    ref = sdf_ref_lib.forward_device(*args)
    got = sdf_loss_forward(*args)
    assert float(got[0][:100].abs().max()) == 0.0 and float(ref[0][:100].abs().max()) == 0.0
    torch.testing.assert_close(got[0], ref[0], rtol=4e-7, atol=1e-7)
    assert torch.equal(ref[2], got[2])


def test_interp_fixture_through_the_operator():
    """tests/golden/sdf_interp.npz (outputs of the reference's host-compiled getValueInterpolated /
    getGradientInterpolated) reproduced by the product operator: identity pose and power-of-two dims make the
    operator's grid coordinates equal the fixture's exactly; eps = 100 puts every in-bounds value on a potential
    branch; potential and gradient follow from the fixture's value / gradient by kernel.cu:158-171."""
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "sdf_interp.npz"))
    grid, pg, val, grad, delta = fx["grid"], fx["pgrid"], fx["value"], fx["grad"], np.float32(fx["delta"])
    d = grid.shape
    lim = np.array([[0, 0, 0, d[0], d[1], d[2], d[0], d[1], d[2], delta]], np.float32)
    eps, pad = np.float32(100.0), np.float32(0.75)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    out = sdf_loss_forward(dev(np.eye(4)[None]), dev(grid[None]), dev(lim), dev(pg), dev([eps]), dev([pad]), dev([0.01]),
                           dev([0.0]))
    pot, g = out[0].cpu().numpy(), out[1].cpu().numpy()
    f32 = np.float32
    neg = val <= 0
    dd = (val - eps).astype(f32)
    inv2, inv1 = f32(1.0) / (f32(2.0) * eps), f32(1.0) / eps
    exp_pot = np.where(neg, (-val.astype(np.float64) + 0.5 * np.float64(eps)).astype(f32),
                       (((inv2 * dd).astype(f32) * dd).astype(f32) * pad).astype(f32))
    exp_g = np.where(neg[:, None], -grad, ((((inv1 * grad).astype(f32)) * dd[:, None]).astype(f32) * pad).astype(f32))
    exp_g = exp_g + f32(0.0)   # R^T v with R = I: fma(1, vx, fma(0, vy, 0*vz)) = vx (+0 normalises -0)
    np.testing.assert_array_equal(pot.view(np.uint32), exp_pot.astype(f32).view(np.uint32))
    np.testing.assert_array_equal((g + f32(0.0)).view(np.uint32), exp_g.astype(f32).view(np.uint32))
