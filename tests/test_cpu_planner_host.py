"""CPU: the host bookkeeping of the Planner mirror that needs no device -- cutting per-trajectory histories and info
lists where the reference's loop would have stopped (omg/planner.py:627-635), the lazily built info dicts, dynamic
waypoint counts, the initial goal choice of grasp_init (omg/planner.py:187-222)."""
import types

import numpy as np

from omg_planner_b200.config import ChompConfig
from omg_planner_b200.core import dynamic_timesteps
from omg_planner_b200.planner import InfoList, Planner, info_from_row


def _planner(cfg):
    p = Planner.__new__(Planner)
    p.cfg = cfg
    return p


def test_assemble_cuts_histories_like_the_reference_loop():
    cfg = ChompConfig(goal_set_proj=False)
    iters, B, n = 6, 3, 4
    rng = np.random.RandomState(0)
    xi0 = rng.rand(B, n, 9)
    h_xi = rng.rand(iters, B, n, 9)
    h_info = rng.rand(iters, B, 16)
    h_info[:, :, 8] = 0
    h_info[2, 1, 8] = 1          # trajectory 1 terminates at iteration index 2
    h_info[0, 2, 8] = 1          # (t = 0 does not count, planner.py:627)
    term = h_info[:, :, 8] > 0
    term[0] = False
    stopped = term.any(0)
    stop = np.where(stopped, term.argmax(0), iters - 1)
    final = rng.rand(B, 16)
    traj = types.SimpleNamespace(set=lambda x: setattr(traj, "data", x), data=None)
    p = _planner(cfg)
    sel = np.arange(iters * B).reshape(iters, B)
    hist = np.concatenate([xi0[None], h_xi], axis=0)         # slot 0 = the initial trajectory (one device buffer)
    sels = p._assemble(traj, True, hist, h_info, stop, stopped, final, h_xi[-1], sel=sel, n_sel=4)
    # not terminated: initial state + every iteration; info list = every iteration + the info-only pass
    assert p.history_trajectories[0].shape[0] == iters + 1 and len(p.info[0]) == iters + 1
    np.testing.assert_array_equal(p.history_trajectories[0][0], xi0[0])
    np.testing.assert_array_equal(p.history_trajectories[0][-1], h_xi[-1, 0])
    assert p.info[0][-1]["cost"] == final[0, 2]
    # terminated at index 2: three iterations ran, the state after the last one is dropped, no extra info
    assert p.history_trajectories[1].shape[0] == 3 and len(p.info[1]) == 3
    np.testing.assert_array_equal(p.history_trajectories[1][-1], h_xi[1, 1])
    assert p.info[1][-1]["terminate"] is True and p.info[1][0]["terminate"] is False
    assert len(p.info[2]) == iters + 1                      # the t = 0 terminate was ignored
    assert sels[0] == [0, 3, 6, 9] and sels[1] == [1, 4, 7]  # selections stop with the plan (and at optim_steps)
    # reference shape for one trajectory: plain lists
    p2 = _planner(cfg)
    p2._assemble(traj, False, hist[:, :1], h_info[:, :1], stop[:1], stopped[:1], final[:1], h_xi[-1, :1])
    assert isinstance(p2.info, list) and isinstance(p2.history_trajectories, list) and traj.data.shape == (n, 9)


def test_info_list_is_lazy_and_sequence_like():
    cfg = ChompConfig()
    rows = np.arange(3 * 16, dtype=float).reshape(3, 16)
    lst = InfoList(cfg, rows, 30, extra=np.full(16, 7.0))
    assert len(lst) == 4 and lst[-1]["cost"] == 7.0 and lst[1]["obs"] == 16.0
    lst[-1]["time"] = 1.5
    assert lst[3]["time"] == 1.5                            # cached dicts persist
    assert [d["smooth"] for d in lst] == [1.0, 17.0, 33.0, 7.0]
    assert lst[0]["standoff_idx"] == 30 - cfg.reach_tail_length
    assert info_from_row(ChompConfig(use_standoff=False), rows[0], 30)["standoff_idx"] == 29
    assert [d["obs"] for d in lst[1:3]] == [16.0, 32.0]


def test_dynamic_timesteps_and_initial_goal_choice():
    cfg = ChompConfig(traj_delta=0.05, traj_min_step=2, traj_max_step=50)
    start = np.zeros((3, 9))
    end = np.stack([np.full(9, 0.001), np.full(9, 0.4), np.full(9, 3.0)])
    np.testing.assert_array_equal(dynamic_timesteps(start, end, cfg), [2, 24, 50])
    # grasp_init: goal_idx = -1 -> argmin(potentials + dist_eps * weighted joint distance); Proj -> nearest goal
    goals = np.stack([np.full(9, v) for v in (1.0, 0.2, 0.5)])
    for alg, pots, want in (("MD", [np.array([0.0, 5.0, 0.0])], 2), ("Proj", [np.array([0.0, 5.0, 0.0])], 1)):
        cfg = ChompConfig(goal_set_proj=True, use_standoff=False, goal_idx=-1, ol_alg=alg)
        target = types.SimpleNamespace(grasps=goals, reach_grasps=[], grasp_potentials=pots)
        env = types.SimpleNamespace(objects=[target], target_idx=0, config=cfg)
        traj = types.SimpleNamespace(start=np.zeros(9), goal_set=[], interpolate_waypoints=lambda: None)
        p = _planner(cfg)
        p.env, p.traj = env, traj
        p.grasp_init(env)
        assert traj.goal_idx == want
        np.testing.assert_array_equal(traj.end, goals[want])


def test_target_without_grasps_leaves_an_empty_goal_set_and_plan_does_not_run():
    """omg/planner.py:192-197 assigns traj.goal_set = target.grasps unconditionally: after a target switch whose IK /
    collision filter yields no goals, the (reused) Trajectory must not keep the previous target's goals, and plan()
    returns the empty info list ("planning not run", planner.py:650-652)."""
    goals = np.stack([np.full(9, v) for v in (1.0, 0.2, 0.5)])
    cfg = ChompConfig(goal_set_proj=True, use_standoff=True, goal_idx=0)
    with_goals = types.SimpleNamespace(grasps=goals, reach_grasps=[], grasp_potentials=[0.0, 0.0, 0.0])
    without = types.SimpleNamespace(grasps=[], reach_grasps=[], grasp_potentials=[])
    env = types.SimpleNamespace(objects=[with_goals, without], target_idx=0, config=cfg)
    calls = []
    traj = types.SimpleNamespace(start=np.zeros(9), goal_set=[], end=np.zeros(9), data=np.zeros((30, 9)),
                                 interpolate_waypoints=lambda: calls.append(1))
    p = _planner(cfg)
    p.env, p.traj = env, traj
    p.grasp_init(env)
    assert len(traj.goal_set) == 3 and traj.goal_idx == 0 and len(calls) == 1
    env.target_idx = 1                       # PlanningScene.update_planner() with a new target, same Trajectory
    p.grasp_init(env)
    assert len(traj.goal_set) == 0 and len(traj.goal_potentials) == 0 and len(calls) == 1
    assert p.plan(traj) == []                # no device work is attempted
