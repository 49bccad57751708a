"""GPU: the raw operator (drop-in for omg_cuda.sdf_loss_forward) through the C ABI vs the C oracle.
Bar: bit-exact (the operator is fp32 with a fixed op order; see csrc/sdf_device.cuh)."""
import numpy as np
import pytest
import torch

from omg_planner_b200 import scene as S
from oracle import sdf_loss_ref as op

pytestmark = pytest.mark.gpu


def _scene_inputs(sc, eps=0.2, seed=0):
    num = len(sc["names"])
    rng = np.random.RandomState(seed)
    from omg_planner_b200.cost import se3_inverse_f32
    poses = np.stack([se3_inverse_f32(sc["pose_mats"][i]) for i in range(num)])
    e = np.full(num, eps, np.float32); e[0] = 0.1
    pad = np.ones(num, np.float32); pad[-1] = 0.5
    clr = np.full(num, 0.01, np.float32); clr[0] = 0.0
    dis = np.zeros(num, np.float32)
    return poses, e, pad, clr, dis, rng


def _points(sc, rng, n):
    """Half the points near objects (inside their grids), half anywhere in the workspace."""
    num = len(sc["names"])
    idx = rng.randint(num, size=n // 2)
    lim = sc["sdf_limits"]
    local = rng.uniform(-0.6, 0.6, (n // 2, 3)) * (lim[idx, 3:6] - lim[idx, 0:3])
    world = np.einsum("nab,nb->na", sc["pose_mats"][idx, :3, :3], local) + sc["pose_mats"][idx, :3, 3]
    anywhere = rng.uniform([-0.2, -0.8, -0.3], [1.2, 0.8, 1.0], (n - n // 2, 3))
    return np.concatenate([world, anywhere]).astype(np.float32)


def _gpu(args):
    import omg_cuda
    t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in args]
    out = omg_cuda.sdf_loss_forward(*t)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


@pytest.mark.parametrize("kw", [dict(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48]),
                                dict(num_objects=10, grid=64, seed=2)])
def test_operator_bit_exact_vs_oracle(kw):
    sc = S.make_scene(**kw)
    poses, e, pad, clr, dis, rng = _scene_inputs(sc)
    pts = _points(sc, rng, 20000)
    ref = op.sdf_loss_forward(poses, sc["sdf_grids"], sc["sdf_limits"], pts, e, pad, clr, dis, return_pin=True)
    got = _gpu((poses, sc["sdf_grids"], sc["sdf_limits"], pts, e, pad, clr, dis))
    assert ref[3] > 2000 and (ref[0] > 0).sum() > 1000, "test scene must exercise the in-bounds paths"
    np.testing.assert_array_equal(got[0], ref[0])
    np.testing.assert_array_equal(got[1], ref[1])
    np.testing.assert_array_equal(got[2], ref[2])


def test_disabled_objects_and_empty_input():
    sc = S.make_scene(num_objects=4, grid=32, seed=5)
    poses, e, pad, clr, dis, rng = _scene_inputs(sc)
    dis[:] = 1
    pts = _points(sc, rng, 1000)
    got = _gpu((poses, sc["sdf_grids"], sc["sdf_limits"], pts, e, pad, clr, dis))
    assert not got[0].any() and not got[1].any() and not got[2].any()
    got = _gpu((poses, sc["sdf_grids"], sc["sdf_limits"], np.zeros((0, 3), np.float32), e, pad, clr, dis))
    assert got[0].shape == (0,) and got[1].shape == (0, 3)


def test_non_axis_aligned_pose_and_large_eps():
    """Rotation with negative trace exercises the second branch of the matrix->quaternion conversion;
    eps >= 1 makes out-of-bounds samples (value 1.0) contribute, as in the reference kernel."""
    sc = S.make_scene(num_objects=3, grid=32, seed=9)
    from scipy.spatial.transform import Rotation
    sc["pose_mats"][1, :3, :3] = Rotation.from_euler("xyz", [3.0, 0.2, -2.9]).as_matrix()
    sc["pose_mats"][0, :3, :3] = Rotation.from_euler("xyz", [0.1, 3.1, 0.3]).as_matrix()
    poses, e, pad, clr, dis, rng = _scene_inputs(sc)
    e[1] = 1.5
    pts = _points(sc, rng, 6000)
    ref = op.sdf_loss_forward(poses, sc["sdf_grids"], sc["sdf_limits"], pts, e, pad, clr, dis)
    got = _gpu((poses, sc["sdf_grids"], sc["sdf_limits"], pts, e, pad, clr, dis))
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)


def test_sdfloss_module_and_input_validation():
    from omg_planner_b200.sdf_matching_loss import SDFLoss
    sc = S.make_scene(num_objects=3, grid=32, seed=4)
    poses, e, pad, clr, dis, rng = _scene_inputs(sc)
    pts = _points(sc, rng, 512)
    args = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in
            (poses, sc["sdf_grids"], sc["sdf_limits"], pts, e, pad, clr, dis)]
    pot, grad, col = SDFLoss()(*args)
    ref = op.sdf_loss_forward(poses, sc["sdf_grids"], sc["sdf_limits"], pts, e, pad, clr, dis)
    np.testing.assert_array_equal(pot.cpu().numpy(), ref[0])
    import omg_cuda
    bad = list(args); bad[3] = bad[3].cpu()
    with pytest.raises(RuntimeError):
        omg_cuda.sdf_loss_forward(*bad)
    bad = list(args); bad[3] = bad[3].t().contiguous().t()  # non-contiguous view
    with pytest.raises(RuntimeError):
        omg_cuda.sdf_loss_forward(*bad)
