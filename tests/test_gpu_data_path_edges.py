"""GPU: edge cases of the entry points either side of the CHOMP loop -- extreme shapes, ragged tails, single
elements, the limits the headers state -- against the oracles."""
import ctypes

import numpy as np
import pytest
import torch

from omg_planner_b200 import _lib
from omg_planner_b200 import core as C
from omg_planner_b200.ik import IkSolver, poses_to_targets
from omg_planner_b200.robot import PandaConstants
from omg_planner_b200.sdf_tools import SignedDensityField
from oracle import chomp_ref as R
from oracle import kdl_ik_ref as K
from oracle import learner_ref as LR
from oracle import sdf_asset_ref as A
from oracle import traj_ref as T

pytestmark = pytest.mark.gpu
vp = ctypes.c_void_p


def _dev(a, dt=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)


def test_interpolate_knot_count_limits_and_single_sample():
    rng = np.random.RandomState(1)
    for K_, n in ((32, 40), (2, 1), (3, 1), (7, 2)):
        wp = rng.uniform(-2, 2, (3, K_, 9))
        for mode in ("cubic", "linear"):
            got = C.interpolate_waypoints_device(_dev(wp), n, mode).cpu().numpy()
            for b in range(3):
                np.testing.assert_allclose(got[b], T.interpolate_waypoints(wp[b], n, mode=mode), rtol=0, atol=1e-11)
    with pytest.raises(RuntimeError):
        C.interpolate_waypoints_device(_dev(rng.uniform(-1, 1, (2, 33, 9))), 10)      # more than 32 knots
    with pytest.raises(RuntimeError):
        C.interpolate_waypoints_device(_dev(rng.uniform(-1, 1, (2, 2, 9))), 10, "quintic")


def test_sdf_pack_tiny_many_planes_and_maximum_object_count():
    rng = np.random.RandomState(2)
    # one voxel
    f = SignedDensityField(np.array([[[0.25]]], np.float32), np.zeros(3), 0.01)
    np.testing.assert_array_equal(C.pack_sdf_grids([f], (1, 1, 1)).cpu().numpy(), [[[[0.25]]]])
    # 64 objects (OMGB_MAX_OBJECTS) x 1100 x-planes = 70400 (object, x) planes: more than one grid.y can hold
    fields, refs = [], []
    for i in range(64):
        shp = (int(rng.randint(1, 1101)), int(rng.randint(1, 7)), int(rng.randint(1, 7)))
        data = rng.uniform(-0.1, 0.3, shp).astype(np.float32)
        fields.append(SignedDensityField(data, np.zeros(3), 0.01)); refs.append(A.FieldRef(data.copy(), np.zeros(3), 0.01))
    fields[0] = SignedDensityField(np.zeros((1100, 6, 6), np.float32), np.zeros(3), 0.01)
    refs[0] = A.FieldRef(np.zeros((1100, 6, 6), np.float32), np.zeros(3), 0.01)
    want, _ = A.combine_sdfs(refs)
    np.testing.assert_array_equal(C.pack_sdf_grids(fields, want.shape[1:]).cpu().numpy(), want)
    with pytest.raises(RuntimeError):
        C.pack_sdf_grids(fields + [fields[0]], want.shape[1:])                         # 65 objects
    with pytest.raises(RuntimeError):
        C.pack_sdf_grids(fields[:2], (4, 4, 4))                                        # padded shape too small


def test_point_sdf_thin_grids_single_point_and_exact_tile_multiples():
    rng = np.random.RandomState(3)
    for npts, margin in ((1, 0.05), (1024, 0.03), (2048, 0.07), (1025, 0.24)):
        pts = rng.uniform([0.3, -0.1, 0.1], [0.34, -0.05, 0.13], (npts, 3))
        f, d64 = C.compute_sdf_from_points(pts, margin=margin, keep_fp64=True)
        want, origin, _ = A.point_sdf(pts, margin=margin)
        assert d64.shape == want.shape
        np.testing.assert_array_equal(d64.cpu().numpy(), want)
        np.testing.assert_array_equal(f.data_torch.cpu().numpy(), want.astype(np.float32))


def _learner_call(prm, xi, coll, goal_set, state, goal_idx, end, rows, shared=0, reach=None, done=None):
    ptr = lambda t: None if t is None else vp(t.data_ptr())
    p, sc, ep, ec, q = state
    _lib.check(_lib.lib().omgb_learner_update(ctypes.byref(prm), xi.shape[0], ptr(xi), ptr(coll), ptr(goal_set), shared,
                                              ptr(reach), ptr(p), ptr(sc), ptr(ep), ptr(ec), ptr(q), ptr(done),
                                              ptr(goal_idx), ptr(end), ptr(rows), None, None,
                                              vp(torch.cuda.current_stream().cuda_stream)), "omgb_learner_update")


@pytest.mark.parametrize("G", [1, 2, 256])
def test_learner_goal_count_limits_shared_goals_proj_and_done(G):
    rng = np.random.RandomState(4)
    B, c = 3, 5
    cfg = R.RefConfig(ol_alg="MD", optim_steps=50)
    refs = [LR.LearnerRef(cfg, G) for _ in range(B)]
    prm = _lib.LearnerParams()
    prm.alg, prm.num_goals, prm.n_waypoints, prm.first_waypoint, prm.constraint_rows = 3, G, 30, 4, c
    prm.normalize_cost, prm.base_obstacle_weight, prm.smoothness_base_weight, prm.dist_eps = 1, 1.0, 0.1, 0.1
    prm.eta = refs[0].eta
    for k in range(5):
        prm.etas[k] = refs[0].etas[k]
    xi = rng.uniform(-1, 1, (B, 30, 9))
    goal_set = rng.uniform(-1, 1, (G, 9))                      # shared by all trajectories
    reach = rng.uniform(-1, 1, (G, c, 9))
    state = [_dev(np.ones((B, G)) / G), _dev(np.zeros((B, G))), _dev(np.ones((B, 5, G)) / G), _dev(np.zeros((B, 5))),
             _dev(np.ones((B, 5)) / 5)]
    goal_idx = torch.zeros(B, dtype=torch.int32, device="cuda")
    end, rows = _dev(np.zeros((B, 9))), _dev(np.zeros((B, c, 9)))
    done = torch.tensor([0, 1, 0], dtype=torch.uint8, device="cuda")   # trajectory 1 keeps its goal
    for t in range(6):
        coll = rng.uniform(0.0, 2.0, (B, G)).astype(np.float32)
        _learner_call(prm, _dev(xi), _dev(coll, torch.float32), _dev(goal_set), state, goal_idx, end, rows, shared=1,
                      reach=_dev(reach), done=done)
        sel = goal_idx.cpu().numpy()
        for b in (0, 2):
            smooth = np.linalg.norm(np.diff(xi[b, 4] - goal_set, axis=-1), axis=-1) ** 2
            cv = 1.0 * coll[b] + 0.1 * 0.1 * smooth
            cv = cv / np.linalg.norm(cv)
            want = refs[b].update(cv)
            assert sel[b] == want
            np.testing.assert_allclose(state[0].cpu().numpy()[b], refs[b].p, rtol=0, atol=1e-6)
            np.testing.assert_array_equal(rows.cpu().numpy()[b], reach[want])
            np.testing.assert_array_equal(end.cpu().numpy()[b], goal_set[want])
        assert sel[1] == 0 and np.abs(state[0].cpu().numpy()[1] - 1.0 / G).max() < 1e-15
    # Proj: nearest goal to the last waypoint, no collision costs needed
    prm.alg = 4
    _learner_call(prm, _dev(xi), None, _dev(goal_set), state, goal_idx, end, rows, shared=1, reach=_dev(reach))
    want = np.argmin(np.linalg.norm(xi[:, -1][:, None] - goal_set[None], axis=-1), axis=1)
    np.testing.assert_array_equal(goal_idx.cpu().numpy(), want)
    prm.num_goals = 257
    with pytest.raises(RuntimeError):
        _learner_call(prm, _dev(xi), None, _dev(goal_set), state, goal_idx, end, rows, shared=1)


def test_ik_single_solves_single_seed_and_unreachable_targets():
    robot = PandaConstants()
    sol = IkSolver(robot.pose_0, robot.joint_lower_limit, robot.joint_upper_limit)
    chain = K.PandaChain(robot.pose_0, robot.joint_lower_limit, robot.joint_upper_limit)
    rng = np.random.RandomState(6)
    q = rng.uniform(chain.lo, chain.hi, (40, 7))
    tg = poses_to_targets(np.stack([chain.fk_hand(v) for v in q]))
    far = tg.copy()
    far[:, :3] += 5.0                                             # five metres away: every solve fails
    seed = np.array([[0.0, -1.285, 0, -2.356, 0.0, 1.571, 0.785]])
    sols, solved, steps = sol.solve_chains(np.concatenate([tg, far])[:, None], seed, want_steps=True)
    assert (solved[40:] == 0).all() and (steps[40:] == 100).all()
    agree = 0
    for p in range(40):
        want, rc, its, raw = chain.ik(tg[p, :3], tg[p, 3:], seed[0])
        agree += (rc >= 0) == (solved[p, 0] == 1)
        if rc >= 0 and solved[p, 0] == 1:
            assert np.abs(sols[p, 0, 0] - want).max() < 1e-3
    assert agree >= 39
    # the seed that already solves the problem: zero Newton steps, the seed comes back bit for bit
    exact = poses_to_targets(np.stack([chain.fk_hand(v) for v in q[:4]]))
    for p in range(4):
        s, ok, st = sol.solve_chains(exact[p][None, None], q[p][None], want_steps=True)
        assert ok[0, 0] == 1 and st[0, 0, 0] <= 1
        assert np.abs(s[0, 0, 0] - q[p]).max() < 1e-6
    assert sol.inverse_kinematics(far[0, :3], far[0, 3:], seed[0]) is None
    got = sol.inverse_kinematics(tg[0, :3], tg[0, 3:], q[0])
    assert got is not None and np.abs(got - q[0]).max() < 1e-5
