"""GPU: edge cases of the fused CHOMP iteration against the oracle -- body-point counts other than 15 (the
32-lanes-per-link-instance kernel shape), disabled objects / "floor" / an attached target (table parameters,
omg/cost.py:303-328), uncheck_finger_collision = -1 (cost.py:350-353), short trajectories (dynamic timesteps,
omg/config.py:96-99), single-object scenes, empty batches."""
import numpy as np
import pytest
import torch

import helpers as H
from omg_planner_b200 import scene as S
from omg_planner_b200.config import ChompConfig
from omg_planner_b200.robot import PandaConstants
from oracle import chomp_ref as R

pytestmark = pytest.mark.gpu
TOL_RAD = 1e-7


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _compare(sc, mode, robot, rref, xi, st, en, tails, iters, cfg_kw=None, n=30):
    cfg = ChompConfig(timesteps=n, **mode, **(cfg_kw or {}))
    rows = H.goal_rows_for(mode, tails, en)
    ref_hist, ref_infos = H.oracle_steps(sc, mode, xi, st, en, rows, iters, robot_ref=rref, cfg_kw=cfg_kw)
    eng = H.engine_for(sc, cfg, robot)
    x = _dev(xi)
    for it in range(iters):
        cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
        out = eng.step(cfg, x, _dev(st), _dev(en), None if rows is None else _dev(rows))
    torch.cuda.synchronize()
    err = np.abs(x.cpu().numpy() - ref_hist[:, iters])[..., :7].max()
    assert err <= TOL_RAD, err
    info = out["info"].cpu().numpy()
    for b in range(xi.shape[0]):
        assert info[b, 3] == ref_infos[b][-1]["collide"]
    return info


@pytest.mark.parametrize("p", [7, 16, 24, 32])
def test_body_point_counts(p):
    base = PandaConstants().collision_points                      # [10,15,3]
    rng = np.random.RandomState(p)
    reps = -(-p // 15)
    pts = np.concatenate([base + (0.004 * rng.randn(*base.shape) if r else 0) for r in range(reps)], axis=1)[:, :p]
    robot = PandaConstants(body_points=pts)
    rref = R.PandaRef(body_points=pts)
    sc = S.make_scene(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
    xi, st, en, tails = S.make_trajectories(5, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=p)
    for name in ("goalset_standoff_topk", "fixed_full"):
        _compare(sc, H.MODES[name], robot, rref, xi, st, en, tails, 3)


def test_disabled_objects_floor_attached_and_soft_fingers():
    robot, rref = PandaConstants(), R.PandaRef()
    sc = S.make_scene(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
    xi, st, en, tails = S.make_trajectories(6, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=13)
    mode = H.MODES["goalset_standoff_topk"]
    base = _compare(sc, mode, robot, rref, xi, st, en, tails, 2)
    # (a) an object named "floor" and one listed in cfg.disable_collision_set contribute nothing
    sc2 = dict(sc); sc2["names"] = list(sc["names"]); sc2["names"][1] = "floor"
    off = _compare(sc2, mode, robot, rref, xi, st, en, tails, 2, cfg_kw=dict(disable_collision_set=[sc["names"][2]]))
    assert (off[:, 12] <= base[:, 12]).all() and off[:, 12].sum() < base[:, 12].sum()      # fewer in-bounds pairs
    # (b) attached target: the last object (table) gets clearance 0, eps 0.05, padding 0.5
    sc3 = dict(sc); sc3["attached"] = True
    _compare(sc3, mode, robot, rref, xi, st, en, tails, 2)
    # (c) uncheck_finger_collision = -1: finger links' potentials x 0.1, never colliding
    _compare(sc, mode, robot, rref, xi, st, en, tails, 2, cfg_kw=dict(uncheck_finger_collision=-1))
    _compare(sc, H.MODES["fixed_full"], robot, rref, xi, st, en, tails, 2, cfg_kw=dict(uncheck_finger_collision=-1))


@pytest.mark.parametrize("n,name", [(3, "fixed_full"), (8, "goalset_standoff_topk"), (5, "goalset_standoff_topk"),
                                    (2, "goalset_single_full"), (50, "fixed_topk")])
def test_short_and_long_trajectories(n, name):
    robot, rref = PandaConstants(), R.PandaRef()
    sc = S.make_scene(num_objects=4, grid=32, seed=5)
    xi, st, en, tails = S.make_trajectories(4, n, robot.joint_lower_limit, robot.joint_upper_limit, seed=n)
    _compare(sc, H.MODES[name], robot, rref, xi, st, en, tails, 3, n=n)


def test_single_object_scene_and_empty_batch():
    robot, rref = PandaConstants(), R.PandaRef()
    sc = S.make_scene(num_objects=1, grid=32, seed=9)
    xi, st, en, tails = S.make_trajectories(3, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=2)
    mode = H.MODES["goalset_standoff_topk"]
    _compare(sc, mode, robot, rref, xi, st, en, tails, 2)
    cfg = ChompConfig(**mode)
    eng = H.engine_for(sc, cfg, robot)
    out = eng.step(cfg, _dev(xi[:0]), _dev(st[:0]), _dev(en[:0]), _dev(tails[:0]))
    assert out["info"].shape == (0, 16)
    out = eng.plan(cfg, _dev(xi[:0]), _dev(st[:0]), _dev(en[:0]), _dev(tails[:0]), iters=3)
    assert out["info"].shape == (0, 16)
