#!/usr/bin/env python
"""Parity of whole goal-set plans with goal switching (device pipeline: omgb_goal_costs -> omgb_learner_update ->
omgb_chomp_plan_step) against the oracle's Planner.plan restatement (oracle/planner_ref.py + learner_ref.py, pinned to
the reference by tests/golden/plan_*.npz and learner_*.npz), on goal sets whose goals are close enough that the
leader changes.

  python tools/parity_report_goalset.py dump  OUT.npz     on the GPU box
  python tools/parity_report_goalset.py check OUT.npz     anywhere (CPU); prints one JSON object
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SCENE = dict(num_objects=8, grid=64, seed=5)
B, G, OPT, EXTRA = 24, 9, 14, 4
ALGS = {"MD": True, "Exp": False, "FTL": True, "FTC": False}     # alg -> use_standoff


def _goals():
    from omg_planner_b200 import scene as S
    from omg_planner_b200.robot import PandaConstants

    robot = PandaConstants()
    return S.make_goal_sets(B, G, robot.joint_lower_limit, robot.joint_upper_limit, seed=31, spread=0.1)


def dump(path):
    import helpers as H
    from omg_planner_b200 import core as C
    from omg_planner_b200 import scene as S
    from omg_planner_b200.config import ChompConfig
    from omg_planner_b200.planner import Planner
    from omg_planner_b200.robot import PandaConstants

    sc = S.make_scene(**SCENE)
    robot = PandaConstants()
    goals, reach = _goals()
    out = {}
    for alg, standoff in ALGS.items():
        cfg = ChompConfig(goal_set_proj=True, use_standoff=standoff, ol_alg=alg, optim_steps=OPT, extra_smooth_steps=EXTRA,
                          pre_terminate=False)
        env = H.make_env(sc, cfg, robot)
        target = env.objects[env.target_idx]
        target.grasps, target.reach_grasps = goals, (reach if standoff else goals)
        traj = C.Trajectory(30, cfg=cfg, start=np.tile(S.START_CONF, (B, 1)), end=goals[:, 0])
        planner = Planner(env, traj)
        out["xi0_" + alg] = np.array(traj.data)
        out["goal0_" + alg] = np.array(traj.goal_idx)
        planner.plan(traj)
        out["hist_" + alg] = np.stack(planner.history_trajectories)
        out["sel_" + alg] = np.array(planner.selected_goals)
    np.savez_compressed(path, **out)
    print("wrote", path)


def _worker(args):
    alg, standoff, b, xi0, g0, goals, reach = args
    from omg_planner_b200 import scene as S
    from oracle import chomp_ref as R
    from oracle import learner_ref as LR
    from oracle import planner_ref as P

    sc = S.make_scene(**SCENE)
    cfg = R.RefConfig(goal_set_proj=True, use_standoff=standoff, top_k_collision=1000, ol_alg=alg, optim_steps=OPT,
                      extra_smooth_steps=EXTRA, pre_terminate=False)
    learner = LR.LearnerRef(cfg, G)
    rows = reach[g0] if standoff else goals[g0][None]
    hist, infos, sel, final = P.plan(R.PandaRef(), sc, cfg, xi0, S.START_CONF, goals[g0], rows, goal_set=goals,
                                     reach_grasps=reach if standoff else goals[:, None, :], goal_idx=g0, learner=learner)
    return alg, b, np.stack(hist), sel


def check(path):
    g = np.load(path)
    goals, reach = _goals()
    jobs = [(alg, standoff, b, g["xi0_" + alg][b], int(g["goal0_" + alg][b]), goals[b], reach[b])
            for alg, standoff in ALGS.items() for b in range(B)]
    report = {"scene": SCENE, "trajectories": B, "goals": G, "iterations": OPT + EXTRA, "algorithms": {}}
    res = {alg: {} for alg in ALGS}
    with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:
        for alg, b, hist, sel in pool.imap_unordered(_worker, jobs, chunksize=2):
            res[alg][b] = (hist, sel)
    for alg in ALGS:
        err, same, switched = [], 0, 0
        for b in range(B):
            hist, sel = res[alg][b]
            err.append(np.abs(g["hist_" + alg][b][:len(hist)] - hist)[..., :7].max())
            same += int(list(g["sel_" + alg][b]) == list(sel))
            switched += int(len(set(sel)) > 1)
        report["algorithms"][alg] = {"max_abs_rad": float(np.max(err)), "fraction_within_1e-4": float((np.array(err) <= 1e-4).mean()),
                                     "selected_goal_sequences_identical": "%d/%d" % (same, B),
                                     "trajectories_whose_goal_switched": switched}
    print(json.dumps(report))


if __name__ == "__main__":
    {"dump": dump, "check": check}[sys.argv[1]](sys.argv[2])
