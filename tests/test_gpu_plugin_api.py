"""GPU: the reference-facing plugin surface (Cost / Optimizer / batch_obstacle_cost) used the way
omg/planner.py:100-101,612-627 and omg/online_learner.py:134 use the reference classes."""
import numpy as np
import pytest
import torch

import helpers as H
from omg_planner_b200 import scene as S
from omg_planner_b200.config import ChompConfig
from omg_planner_b200.cost import Cost
from omg_planner_b200.optimizer import Optimizer
from omg_planner_b200.robot import PandaConstants
from oracle import chomp_ref as R

pytestmark = pytest.mark.gpu


def test_optimizer_optimize_matches_oracle_loop():
    mode = H.MODES["goalset_standoff_topk"]
    sc = S.make_scene(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
    cfg = ChompConfig(**mode)
    robot = PandaConstants()
    env = H.make_env(sc, cfg, robot)
    xi, st, en, tails = S.make_trajectories(2, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=4)
    cost = Cost(env)
    for b in range(2):
        optim = Optimizer(env, cost)
        env.objects[env.target_idx].reach_grasps = [tails[b]]
        cost.target_obj = env.objects[env.target_idx]
        traj = H.FakeTrajectory(xi[b], st[b], en[b], goal_set=[en[b]], goal_idx=0)
        ref = R.ChompRef(R.PandaRef(), sc, R.RefConfig(**mode), xi[b], st[b], en[b], tails[b])
        for t in range(12):
            info = optim.optimize(traj, force_update=True)
            rinfo = ref.step()
            assert np.abs(traj.data - ref.xi)[:, :7].max() <= 1e-7
            for key in ("obs", "smooth", "cost", "collide", "reach", "grad", "weighted_obs_grad", "weighted_smooth_grad"):
                slack = rinfo["tie_slack"] * (1 + 1e-9) if key in ("obs", "cost") else 0.0
                assert abs(info[key] - rinfo[key]) <= 1e-6 * max(1.0, abs(rinfo[key])) + slack, key
            for key in ("terminate", "violate_limit", "execute", "failure_terminate"):
                assert bool(info[key]) == bool(rinfo[key]), key
            np.testing.assert_allclose(info["gradient"], rinfo["gradient"], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(info["cost_traj"], rinfo["cost_traj"], rtol=1e-6, atol=1e-6 + rinfo["tie_slack"])
            assert info["standoff_idx"] == rinfo["standoff_idx"]
        assert abs(cfg.smoothness_weight - 0.1 * 1.02 ** 12) < 1e-15   # schedules are written back into cfg
        final = optim.optimize(traj, info_only=True)
        assert "terminate" in final and "text" in final


def test_compute_total_loss_and_batched_trajectory_object():
    mode = H.MODES["fixed_full"]
    sc = S.make_scene(num_objects=5, grid=40, seed=6)
    cfg = ChompConfig(**mode)
    robot = PandaConstants()
    env = H.make_env(sc, cfg, robot)
    xi, st, en, tails = S.make_trajectories(3, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=9)
    cost = Cost(env)
    traj = H.FakeTrajectory(xi[0], st[0], en[0])
    c, g, info = cost.compute_total_loss(traj)
    rcfg = R.RefConfig(**mode)
    rc, rg, rinfo = R.total_cost(R.PandaRef(), sc, rcfg, xi[0], st[0], en[0], None)
    assert abs(c - rc) <= 1e-6 * max(1, abs(rc))
    np.testing.assert_allclose(g, rg, rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(traj.data, xi[0])
    btraj = H.FakeTrajectory(xi, st, en)
    costs, grads, infos = cost.compute_total_loss(btraj)
    assert grads.shape == (3, 30, 9) and abs(costs[0] - c) == 0


def test_batch_obstacle_cost_with_and_without_arc_length():
    sc = S.make_scene(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
    cfg = ChompConfig()
    robot = PandaConstants()
    env = H.make_env(sc, cfg, robot)
    cost = Cost(env)
    rng = np.random.RandomState(0)
    lo, hi = robot.joint_lower_limit[0], robot.joint_upper_limit[0]
    goals = rng.uniform(lo, hi, (7, 9))
    pot, grad, vis, col = cost.batch_obstacle_cost(goals, special_check_id=0, uncheck_finger_collision=-1)
    rp, rg, rc = R.batch_obstacle_cost(R.PandaRef(), sc, R.RefConfig(), goals, -1, -1)
    np.testing.assert_array_equal(pot.cpu().numpy(), rp)
    np.testing.assert_array_equal(grad.cpu().numpy(), rg)
    np.testing.assert_array_equal(col.cpu().numpy(), rc)
    # Learner.cost_vector usage (online_learner.py:128-141): G goals x n' interpolated configurations
    start = S.START_CONF
    n = 13
    t = np.linspace(0, 1, n + 2)[1:-1]
    traj = (start[None, None] + (goals - start)[:, None] * t[None, :, None]).reshape(-1, 9)
    pot, _, _, _ = cost.batch_obstacle_cost(traj, arc_length=n, special_check_id=0, uncheck_finger_collision=0,
                                            start=start, end=goals)
    rp, _, _ = R.batch_obstacle_cost(R.PandaRef(), sc, R.RefConfig(), traj, n, 0, start)
    np.testing.assert_allclose(pot.cpu().numpy(), rp, rtol=2e-5, atol=1e-6)


def test_batched_optimize_hands_out_arrays_that_stay_put():
    """Optimizer.optimize replaces traj.data by a new array every call (omg/core.py:43-57) and the caller may keep the
    old ones (Planner.history_trajectories).  Here the new array is a view of a pinned block that the next call copies
    from directly and that is recycled once dropped: arrays still held must never change, the batched path must
    equal per-trajectory calls, and a caller that edits traj.data in place (or swaps in its own array) must be seen."""
    mode = H.MODES["goalset_standoff_topk"]
    sc = S.make_scene(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
    robot = PandaConstants()
    xi, st, en, tails = S.make_trajectories(5, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=4)

    def run(batched_ix, edit):
        cfg = ChompConfig(**mode)
        env = H.make_env(sc, cfg, robot)
        cost = Cost(env)
        optim = Optimizer(env, cost)
        ix = batched_ix
        env.objects[env.target_idx].reach_grasps = tails[ix][:, None]
        cost.target_obj = env.objects[env.target_idx]
        traj = H.FakeTrajectory(xi[ix].copy(), st[ix], en[ix], goal_set=en[ix][:, None], goal_idx=np.zeros(len(ix), dtype=int))
        held, snaps = [], []
        for t in range(8):
            optim.optimize(traj, force_update=True)
            held.append(traj.data)
            snaps.append(traj.data.copy())
            if edit and t == 3:
                traj.data[:, 5, 2] += 0.01            # in place, in the pinned block
            if edit and t == 5:
                traj.data = traj.data * 1.0           # the caller's own (pageable) array
        for a, b in zip(held, snaps):
            if not (edit and a is held[3]):
                np.testing.assert_array_equal(a, b)
        return traj.data.copy()

    all5 = run(np.arange(5), edit=False)
    for b in range(5):
        np.testing.assert_array_equal(run(np.array([b]), edit=False)[0], all5[b])
    edited = run(np.arange(5), edit=True)
    assert np.abs(edited - all5).max() > 1e-6          # the in-place edit went through the next iterations
