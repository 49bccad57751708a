"""GPU: goal scoring.  omgb_goal_costs against the oracle's restatement of Learner.cost_vector's device
half, the Learner mirror against the fixtures recorded from the reference's own Learner, and the batched Learner
against per-trajectory runs.

Tolerance: the reference sums fp32 potentials with torch reductions (order implementation-defined); the kernel
accumulates the same fp32 terms in fp64.  Bar: 2e-5 relative on the cost vector (fp32 summation noise over
<= 4500 terms), exact agreement of the selected goals on the fixtures."""
import glob
import os

import numpy as np
import pytest
import torch

import helpers as H
from omg_planner_b200 import scene as S
from omg_planner_b200.config import ChompConfig
from omg_planner_b200.cost import Cost
from omg_planner_b200.online_learner import Learner, bregman_projection_rows
from omg_planner_b200.optimizer import Optimizer
from omg_planner_b200.robot import PandaConstants
from oracle import chomp_ref as R
from oracle import learner_ref as LR

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "learner_*.npz")))


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


# line lengths 30 / 19 (one goal per CTA), 12 / 7 (2 / 4 goals per CTA, the last CTA ragged), 10 (3 per CTA), 1 (all 9)
@pytest.mark.parametrize("first,shared", [(0, False), (11, False), (29, True), (20, True), (18, False), (23, True)])
def test_goal_costs_vs_oracle(first, shared):
    sc = S.make_scene(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
    cfg = ChompConfig(goal_set_proj=True, use_standoff=True)
    robot = PandaConstants()
    B, G, n = 4, 9, 30
    xi, st, en, tails = S.make_trajectories(B, n, robot.joint_lower_limit, robot.joint_upper_limit, seed=3)
    goals, _ = S.make_goal_sets(B, G, robot.joint_lower_limit, robot.joint_upper_limit, seed=8)
    eng = H.engine_for(sc, cfg, robot)
    g_in = goals[0] if shared else goals
    out = eng.goal_costs(_dev(xi), first, _dev(g_in), cfg.time_interval, 0).cpu().numpy()
    assert out.shape == (B, G) and out.dtype == np.float32
    rcfg = R.RefConfig()
    ref = np.stack([LR.collision_costs(R.PandaRef(), sc, rcfg, xi[b, first], goals[0] if shared else goals[b], n - first)
                    for b in range(B)])
    assert ref.max() > 0
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=1e-6)
    # lower-bound culling is exact here too
    eng.set_options(use_lower_bound=0)
    out2 = eng.goal_costs(_dev(xi), first, _dev(g_in), cfg.time_interval, 0).cpu().numpy()
    np.testing.assert_array_equal(out, out2)


def test_goal_costs_edge_cases():
    sc = S.make_scene(num_objects=3, grid=32, seed=2)
    cfg = ChompConfig()
    robot = PandaConstants()
    eng = H.engine_for(sc, cfg, robot)
    xi, st, en, tails = S.make_trajectories(2, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=1)
    goals, _ = S.make_goal_sets(2, 3, robot.joint_lower_limit, robot.joint_upper_limit, seed=1)
    assert eng.goal_costs(_dev(xi), 0, _dev(goals[:, :0]), 0.1, 0).shape == (2, 0)      # no goals
    with pytest.raises(RuntimeError):
        eng.goal_costs(_dev(xi), 30, _dev(goals), 0.1, 0)                                # first waypoint out of range
    # uncheck_finger_collision = -1 scales the finger links by 0.1 (omg/cost.py:350-353)
    a = eng.goal_costs(_dev(xi), 5, _dev(goals), 0.1, 0).cpu().numpy()
    b = eng.goal_costs(_dev(xi), 5, _dev(goals), 0.1, -1).cpu().numpy()
    assert (b <= a + 1e-6).all()


class _Traj(H.FakeTrajectory):
    def interpolate_waypoints(self, waypoints=None, mode="cubic"):
        n = 30
        if np.ndim(self.end) == 2:
            self.data = np.stack([S.clamped_cubic(np.asarray(self.start, dtype=np.float64).reshape(-1, 9)[min(b, np.asarray(self.start).reshape(-1, 9).shape[0] - 1)],
                                                  self.end[b], n) for b in range(self.end.shape[0])])
        else:
            self.data = S.clamped_cubic(self.start, self.end, n)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_learner_and_optimizer_track_reference_fixture(path):
    """Planner.plan's interleave (learner.update_goal(); optim.optimize(traj, force_update=True)) through the plugin
    surface vs the reference's recorded cost vectors, goal distributions, selected goals and trajectories."""
    g = np.load(path)
    alg, standoff = str(g["alg"]), bool(int(g["use_standoff"]))
    sc = S.make_scene(**eval(str(g["scene_args"])))
    cfg = ChompConfig(goal_set_proj=True, use_standoff=standoff, top_k_collision=1000, ol_alg=alg)
    robot = PandaConstants(body_points=g["body_points"])
    env = H.make_env(sc, cfg, robot)
    goals, reach, start = g["goals"], g["reach"], g["start"]
    iters = g["history"].shape[1] - 1
    for b in range(goals.shape[0]):
        target = env.objects[env.target_idx]
        target.reach_grasps = reach[b] if standoff else goals[b]
        cost = Cost(env)
        optim = Optimizer(env, cost)
        traj = _Traj(np.zeros((30, 9)), start, goals[b, 0], goal_set=list(goals[b]), goal_idx=0)
        traj.interpolate_waypoints()
        learner = Learner(env, traj, cost)
        np.testing.assert_allclose(learner.cost_vector(), g["cost_vectors"][b, 0], rtol=2e-5, atol=1e-7)
        assert traj.goal_idx == g["selected"][b, 0]
        for it in range(iters):
            learner.update_goal()
            np.testing.assert_allclose(learner.cost_vector(), g["cost_vectors"][b, it + 1], rtol=2e-5, atol=1e-7)
            np.testing.assert_allclose(learner.p, g["p"][b, it + 1], rtol=2e-4, atol=1e-7)
            assert traj.goal_idx == g["selected"][b, it + 1]
            optim.optimize(traj, force_update=True)
            assert np.abs(traj.data - g["history"][b, it + 1])[:, :7].max() <= 1e-7


def test_batched_learner_equals_per_trajectory_learners():
    sc = S.make_scene(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
    robot = PandaConstants()
    B, G = 5, 6
    goals, reach = S.make_goal_sets(B, G, robot.joint_lower_limit, robot.joint_upper_limit, seed=21)
    start = S.START_CONF.copy()
    results = []
    for batched in (True, False):
        cfg = ChompConfig(goal_set_proj=True, use_standoff=True, ol_alg="MD")
        env = H.make_env(sc, cfg, robot)
        target = env.objects[env.target_idx]
        sel, hist = [], []
        groups = [list(range(B))] if batched else [[b] for b in range(B)]
        for grp in groups:
            cost = Cost(env)
            optim = Optimizer(env, cost)
            if batched:
                target.reach_grasps = reach
                traj = _Traj(np.zeros((B, 30, 9)), np.tile(start, (B, 1)), goals[:, 0], goal_set=goals, goal_idx=np.zeros(B, int))
            else:
                b = grp[0]
                target.reach_grasps = reach[b]
                traj = _Traj(np.zeros((30, 9)), start, goals[b, 0], goal_set=list(goals[b]), goal_idx=0)
            traj.interpolate_waypoints()
            learner = Learner(env, traj, cost)
            s = [np.atleast_1d(traj.goal_idx).copy()]
            for it in range(5):
                learner.update_goal()
                s.append(np.atleast_1d(traj.goal_idx).copy())
                optim.optimize(traj, force_update=True)
            sel.append(np.stack(s, 1)); hist.append(np.asarray(traj.data).reshape(-1, 30, 9))
        results.append((np.concatenate(sel), np.concatenate(hist)))
    np.testing.assert_array_equal(results[0][0], results[1][0])
    np.testing.assert_array_equal(results[0][1], results[1][1])
