"""GPU: the learner's goal re-weighting on the device (omgb_learner_update) and the device-resident goal-set plan
(omgb_goal_costs -> omgb_learner_update -> omgb_chomp_plan_step), against the oracle's LearnerRef (pinned to the
reference's Learner by tests/golden/learner_*.npz) and against the host-learner path of the same Planner.
Tolerance: the distributions come out of a bisection that stops at 1e-6 and of exp/log that differ from numpy's in the
last bit -> 1e-6 absolute on p; selections must be identical."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

import helpers as H
from omg_planner_b200 import _lib
from omg_planner_b200 import core as C
from omg_planner_b200 import scene as S
from omg_planner_b200.config import ChompConfig
from omg_planner_b200.planner import Planner
from omg_planner_b200.robot import PandaConstants
from oracle import chomp_ref as R
from oracle import learner_ref as LR

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "learner_*.npz")))
vp = ctypes.c_void_p


def _update(prm, xi, coll, goal_set, state, goal_idx, end, rows, cv_out):
    p, sc, ep, ec, q = state
    ptr = lambda t: None if t is None else vp(t.data_ptr())
    _lib.check(_lib.lib().omgb_learner_update(ctypes.byref(prm), xi.shape[0], ptr(xi), ptr(coll), ptr(goal_set), 0, None,
                                              ptr(p), ptr(sc), ptr(ep), ptr(ec), ptr(q), None, ptr(goal_idx), ptr(end),
                                              ptr(rows), ptr(cv_out), None, vp(torch.cuda.current_stream().cuda_stream)),
               "omgb_learner_update")


@pytest.mark.parametrize("alg", ["MD", "Exp", "FTL", "FTC"])
def test_update_rules_on_cost_streams_vs_oracle(alg):
    """The update rule alone: the kernel is fed collision costs that ARE the cost vector (weights 1 / 0, no
    normalisation), 20 iterations, 16 trajectories with different streams and goal counts up to 100."""
    rng = np.random.RandomState(3)
    for G in (7, 33, 100):
        B, T = 16, 20
        cfg = R.RefConfig(ol_alg=alg, optim_steps=50)
        refs = [LR.LearnerRef(cfg, G) for _ in range(B)]
        prm = _lib.LearnerParams()
        prm.alg, prm.num_goals, prm.n_waypoints, prm.first_waypoint, prm.constraint_rows = _lib.LEARNER_ALGS[alg], G, 30, 0, 1
        prm.normalize_cost, prm.base_obstacle_weight, prm.smoothness_base_weight, prm.dist_eps = 0, 1.0, 0.0, 0.1
        prm.eta = refs[0].eta
        for k in range(5):
            prm.etas[k] = refs[0].etas[k]
        dev = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)
        xi = dev(np.zeros((B, 30, 9)))
        goal_set = dev(rng.uniform(-1, 1, (B, G, 9)))
        state = [dev(np.ones((B, G)) / G), dev(np.zeros((B, G))), dev(np.ones((B, 5, G)) / G), dev(np.zeros((B, 5))),
                 dev(np.ones((B, 5)) / 5)]
        goal_idx = torch.zeros(B, dtype=torch.int32, device="cuda")
        end, rows = dev(np.zeros((B, 9))), dev(np.zeros((B, 1, 9)))
        cv_out = dev(np.zeros((B, G)))
        worst = 0.0
        for t in range(T):
            cv = rng.uniform(0.02, 0.5, (B, G)).astype(np.float32)
            if t % 5 == 4:
                cv[:, rng.randint(G)] *= 0.05   # a new leader now and then
            _update(prm, xi, dev(cv, torch.float32), goal_set, state, goal_idx, end, rows, cv_out)
            np.testing.assert_array_equal(cv_out.cpu().numpy(), cv.astype(np.float64))
            p = state[0].cpu().numpy()
            sel = goal_idx.cpu().numpy()
            for b in range(B):
                want = refs[b].update(cv[b].astype(np.float64))
                worst = max(worst, np.abs(p[b] - refs[b].p).max())
                assert sel[b] == want, (alg, G, t, b)
            np.testing.assert_array_equal(end.cpu().numpy(), goal_set.cpu().numpy()[np.arange(B), sel])
        print(alg, "G=%d worst |p - oracle| = %.2e" % (G, worst))
        assert worst < 1e-6
        if alg == "MD":
            np.testing.assert_allclose(state[4].cpu().numpy(), np.stack([r.q for r in refs]), rtol=0, atol=1e-6)


def _planner(sc, robot, goals, reach, cfg):
    env = H.make_env(sc, cfg, robot)
    target = env.objects[env.target_idx]
    target.grasps = goals
    target.reach_grasps = reach if cfg.use_standoff else goals
    start = np.tile(S.START_CONF, (goals.shape[0], 1)) if goals.ndim == 3 else S.START_CONF
    traj = C.Trajectory(30, cfg=cfg, start=start, end=goals[..., 0, :])
    return Planner(env, traj), traj


@pytest.mark.parametrize("alg,standoff", [("MD", True), ("Exp", False), ("FTL", True)])
def test_device_plan_equals_host_learner_plan(alg, standoff):
    """A batch of 12 goal-set plans with close goals (the leader changes): device pipeline vs the host-learner path."""
    sc = S.make_scene(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
    robot = PandaConstants()
    B, G = 12, 9
    goals, reach = S.make_goal_sets(B, G, robot.joint_lower_limit, robot.joint_upper_limit, seed=21, spread=0.1)
    out = {}
    for host in (True, False):
        cfg = ChompConfig(goal_set_proj=True, use_standoff=standoff, ol_alg=alg, optim_steps=14, extra_smooth_steps=4,
                          pre_terminate=False, host_learner=host)
        planner, traj = _planner(sc, robot, goals, reach, cfg)
        planner.plan(traj)
        out[host] = (np.array(planner.selected_goals), np.stack(planner.history_trajectories), np.array(traj.data),
                     np.array(planner.learner.p), [[i["cost"] for i in lst] for lst in planner.info])
    np.testing.assert_array_equal(out[True][0], out[False][0])
    assert np.abs(out[True][1] - out[False][1]).max() < 1e-9
    assert np.abs(out[True][2] - out[False][2]).max() < 1e-9
    np.testing.assert_allclose(out[True][3], out[False][3], rtol=0, atol=1e-6)
    np.testing.assert_allclose(np.array(out[True][4]), np.array(out[False][4]), rtol=1e-9, atol=1e-9)
    print(alg, "goals selected per trajectory:", [sorted(set(r.tolist())) for r in out[False][0]])


def test_device_plan_honours_cfg_timeout():
    """cfg.timeout (omg/planner.py:629) inside omgb_chomp_plan_goalset: with a zero budget the loop stops at its first
    clock check (iteration 8); what ran is the head of the untimed plan, bit for bit, and the bookkeeping (Optimizer.step,
    Learner.t, selected goals, the closing info-only call) is that of a loop that ran 8 iterations."""
    sc = S.make_scene(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
    robot = PandaConstants()
    B, G = 5, 9
    goals, reach = S.make_goal_sets(B, G, robot.joint_lower_limit, robot.joint_upper_limit, seed=21, spread=0.1)
    out = {}
    for timeout in (-1, 0.0):
        cfg = ChompConfig(goal_set_proj=True, use_standoff=True, ol_alg="MD", optim_steps=14, extra_smooth_steps=4,
                          pre_terminate=False)
        cfg.timeout = timeout
        planner, traj = _planner(sc, robot, goals, reach, cfg)
        t0, step0 = planner.learner.t, planner.optim.step
        planner.plan(traj)
        out[timeout] = (planner.history_trajectories, planner.selected_goals, planner.info,
                        planner.optim.step - step0, planner.learner.t - t0)
    full, cut = out[-1], out[0.0]
    assert full[3] == 18 + 1 and cut[3] == 8 + 1            # one update() per iteration + the info-only call
    assert full[4] == 14 and cut[4] == 8
    for b in range(B):
        assert len(full[0][b]) == 19 and len(cut[0][b]) == 9
        np.testing.assert_array_equal(np.asarray(cut[0][b]), np.asarray(full[0][b])[:9])
        assert list(cut[1][b]) == list(full[1][b])[:8]
        assert len(cut[2][b]) == 9
        for k in range(8):
            assert cut[2][b][k]["cost"] == full[2][b][k]["cost"]


def test_device_cost_vector_matches_learner_mirror():
    """The cost vector the kernel builds == Learner.cost_vector (pinned to the reference by the learner fixtures)."""
    g = np.load(GOLDEN[0])
    sc = S.make_scene(**eval(str(g["scene_args"])))
    robot = PandaConstants(body_points=g["body_points"])
    standoff = bool(int(g["use_standoff"]))
    cfg = ChompConfig(goal_set_proj=True, use_standoff=standoff, ol_alg=str(g["alg"]))
    goals, reach = g["goals"], g["reach"]
    planner, traj = _planner(sc, robot, goals, reach, cfg)
    from omg_planner_b200.online_learner import DeviceLearnerState

    lrn = planner.learner
    want = lrn.cost_vector()            # t = 0
    cost = planner.cost
    xi, start, end, rows, _ = cost._traj_tensors(traj)
    st = DeviceLearnerState(lrn, xi.device)
    lrn.t = -1.0                        # .update() advances t first
    cv = torch.zeros((xi.shape[0], goals.shape[1]), dtype=torch.float64, device=xi.device)
    st.update(cost.engine, xi, end, rows, cost_vector=cv)
    np.testing.assert_allclose(cv.cpu().numpy(), want, rtol=1e-12, atol=1e-14)
