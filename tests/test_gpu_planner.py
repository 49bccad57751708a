"""GPU: the Planner mirror (omg_planner_b200/planner.py) against fixtures recorded from the reference's own
Planner.plan (tools/make_golden_plan.py): history_trajectories, the info list, selected goals and the final
trajectory, one trajectory at a time (the reference's shape) and as one batch (one persistent launch)."""
import glob
import os

import numpy as np
import pytest

import helpers as H
from omg_planner_b200 import core as C
from omg_planner_b200 import scene as S
from omg_planner_b200.config import ChompConfig
from omg_planner_b200.planner import Planner
from omg_planner_b200.robot import PandaConstants

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "plan_*.npz")))
TOL_RAD = 1e-7
COLS = {"obs": "obs", "smooth": "smooth", "cost": "cost", "collide": "collide", "reach": "reach", "grad": "grad",
        "weighted_obs_grad": "weighted_obs_grad", "weighted_smooth_grad": "weighted_smooth_grad"}


def _cfg(g):
    return ChompConfig(goal_set_proj=bool(g["goal_set_proj"]), use_standoff=bool(g["use_standoff"]),
                       ol_alg=str(g["ol_alg"]), optim_steps=int(g["optim_steps"]),
                       extra_smooth_steps=int(g["extra_smooth_steps"]), pre_terminate=bool(g["pre_terminate"]),
                       top_k_collision=1000)


def _setup(g, sc, robot, sel):
    """Planner for the trajectories `sel` of the fixture (an int: reference shape; a list: batched)."""
    cfg = _cfg(g)
    env = H.make_env(sc, cfg, robot)
    target = env.objects[env.target_idx]
    learner_on = cfg.goal_set_proj and cfg.ol_alg not in ("Baseline", "Proj")
    one = np.isscalar(sel)
    idx = sel if one else np.asarray(sel)
    if learner_on:
        target.grasps = g["goals"][idx]
        target.reach_grasps = g["reach"][idx] if cfg.use_standoff else g["goals"][idx]
        end = g["goals"][idx][..., 0, :]
    else:
        end = g["end"][idx]
        if cfg.goal_set_proj:
            target.grasps = g["end"][idx][..., None, :]
            target.reach_grasps = g["tails"][idx][..., None, :, :]
    traj = C.Trajectory(30, cfg=cfg, start=g["start"][idx], end=end)
    return Planner(env, traj), traj, cfg


def _check(g, b, info, hist, selected, final):
    keys, fkeys = [str(k) for k in g["info_keys"]], [str(k) for k in g["flag_keys"]]
    assert len(hist) == int(g["history_len"][b]), (len(hist), int(g["history_len"][b]))
    assert len(info) == int(g["info_len"][b])
    err = np.abs(np.stack(hist) - g["history"][b, :len(hist)])[..., :7].max()
    assert err <= TOL_RAD, err
    assert np.abs(final - g["final"][b])[..., :7].max() <= TOL_RAD
    assert list(selected) == g["selected"][b, :int(g["selected_len"][b])].tolist()
    for k, i in enumerate(info):
        for c, key in enumerate(keys):
            want = g["infos"][b, k, c]
            slack = g["tie_slack"][b, k] * (1 + 1e-9) if key in ("obs", "cost") else 0.0
            assert abs(float(i[key]) - want) <= 1e-6 * max(1.0, abs(want)) + slack, (key, b, k, float(i[key]), want)
        for c, key in enumerate(fkeys):
            assert int(bool(i[key])) == int(g["flags"][b, k, c]), (key, b, k)
    assert "time" in info[-1]
    return err


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[5:-4] for p in GOLDEN])
def test_planner_reproduces_reference_plan(path):
    g = np.load(path)
    sc = S.make_scene(**eval(str(g["scene_args"])))
    robot = PandaConstants(body_points=g["body_points"])
    B = g["xi0"].shape[0]
    worst = 0.0
    for b in range(B):
        planner, traj, cfg = _setup(g, sc, robot, b)
        assert np.abs(traj.data - g["xi0"][b]).max() <= 1e-12        # initial state = the reference's
        info = planner.plan(traj)
        worst = max(worst, _check(g, b, info, planner.history_trajectories, planner.selected_goals, traj.data))
    print("worst |history - reference| = %.2e rad" % worst)
    # the same trajectories as ONE batch
    planner, traj, cfg = _setup(g, sc, robot, list(range(B)))
    infos = planner.plan(traj)
    for b in range(B):
        sel = planner.selected_goals[b] if planner.selected_goals else []
        _check(g, b, infos[b], planner.history_trajectories[b], sel, traj.data[b])


def test_fused_plan_history_equals_per_iteration_launches():
    """omgb_chomp_plan_history's records == what a host loop over omgb_chomp_step sees, bit for bit, including
    trajectories frozen by stop_on_terminate."""
    import torch

    mode = H.MODES["fixed_topk"]
    sc = S.make_scene(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
    cfg = ChompConfig(**mode)
    robot = PandaConstants()
    xi, st, en, tails = S.make_trajectories(24, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=5)
    eng = H.engine_for(sc, cfg, robot)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x1, s, e = dev(xi), dev(st), dev(en)
    iters = 30
    out = eng.plan(cfg, x1, s, e, None, iters=iters, stop_on_terminate=True, history=True)
    hx, hi = out["hist_xi"].cpu().numpy(), out["hist_info"].cpu().numpy()
    x2 = dev(xi)
    done = np.zeros(24, bool)
    for it in range(iters):
        cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
        act = torch.from_numpy((~done).astype(np.uint8)).cuda()
        info = eng.step(cfg, x2, s, e, None, active=act, update=1)["info"].cpu().numpy()
        live = ~done
        np.testing.assert_array_equal(hx[it][live], x2.cpu().numpy()[live])
        np.testing.assert_array_equal(hi[it][live], info[live])
        if it > 0:
            np.testing.assert_array_equal(hx[it][done], hx[it - 1][done])
            done |= live & (info[:, 8] > 0)
    np.testing.assert_array_equal(x1.cpu().numpy(), x2.cpu().numpy())
    assert done.any() and not done.all()
