#!/usr/bin/env python
"""Generate tests/golden/learner_*.npz by running the UNMODIFIED reference Learner + Cost + Optimizer
(omg/online_learner.py, omg/cost.py, omg/optimizer.py from /root/reference under the stubs of
tools/ref_harness.py) in the interleave of Planner.plan (omg/planner.py:612-621: learner.update_goal(), then
optim.optimize(traj, force_update=True)) on synthetic scenes with synthetic goal sets.

BUILD-CONTAINER ONLY (needs /root/reference).  The committed .npz files are what travels.
Recorded per iteration: the cost vector the Learner computed, its goal distribution p, the selected goal, and the
trajectory after the CHOMP step.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as H  # noqa: E402
from omg_planner_b200 import scene as S  # noqa: E402
from oracle import chomp_ref as R  # noqa: E402

CASES = {
    "md_standoff": dict(ol_alg="MD", use_standoff=True, spread=0.6),
    "md_close_goals": dict(ol_alg="MD", use_standoff=True, spread=0.1),
    "exp_single": dict(ol_alg="Exp", use_standoff=False, spread=0.6),
    "ftl_standoff": dict(ol_alg="FTL", use_standoff=True, spread=0.1),
    "ftc_single": dict(ol_alg="FTC", use_standoff=False, spread=0.1),
}
SCENE_ARGS = dict(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
N_TRAJ, N_GOALS, N_WPT, N_ITER = 3, 7, 30, 14


def main():
    ns = H.load_reference()
    cfg = ns.cfg
    cfg.timeout = -1
    cfg.report_cost = False
    cfg.report_time = False
    sc = S.make_scene(**SCENE_ARGS)
    robot = R.PandaRef()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, case in CASES.items():
        cfg.goal_set_proj = True
        cfg.top_k_collision = 1000
        cfg.use_standoff = case["use_standoff"]
        cfg.ol_alg = case["ol_alg"]
        cfg.timesteps = N_WPT
        cfg.get_global_param(N_WPT)
        env = H.make_ref_env(ns, sc, robot.body_points)
        goals, reach = S.make_goal_sets(N_TRAJ, N_GOALS, robot.lower, robot.upper, seed=3, spread=case["spread"])
        start = S.START_CONF.copy()
        cvs = np.zeros((N_TRAJ, N_ITER + 1, N_GOALS)); ps = np.zeros((N_TRAJ, N_ITER + 1, N_GOALS))
        sel = np.zeros((N_TRAJ, N_ITER + 1), np.int64)
        hist = np.zeros((N_TRAJ, N_ITER + 1, N_WPT, 9))
        for b in range(N_TRAJ):
            target = env.objects[env.target_idx]
            target.reach_grasps = reach[b] if cfg.use_standoff else goals[b]
            cost = ns.cost.Cost(env)
            opt = ns.optimizer.Optimizer(env, cost)
            traj = H.RefTrajectory(ns, np.zeros((N_WPT, 9)), start, goals[b, 0], goal_set=list(goals[b]), goal_idx=0)
            traj.interpolate_waypoints()
            learner = ns.online_learner.Learner(env, traj, cost)   # picks the initial goal (online_learner.py:95-102)
            cvs[b, 0] = learner.cost_vector()
            ps[b, 0] = learner.p
            sel[b, 0] = traj.goal_idx
            hist[b, 0] = traj.data
            for it in range(N_ITER):
                learner.update_goal()
                # the cost vector update_goal just used (t was incremented first; recomputing is deterministic)
                cvs[b, it + 1] = learner.cost_vector()
                ps[b, it + 1] = learner.p
                sel[b, it + 1] = traj.goal_idx
                opt.optimize(traj, force_update=True)
                hist[b, it + 1] = traj.data
        # the update rule alone on a seeded stream of synthetic cost vectors (so that the leader changes often)
        rng = np.random.RandomState(17)
        syn_cv = rng.uniform(0.05, 1.0, (20, N_GOALS))
        syn_cv /= np.linalg.norm(syn_cv, axis=1, keepdims=True)
        traj = H.RefTrajectory(ns, np.zeros((N_WPT, 9)), start, goals[0, 0], goal_set=list(goals[0]), goal_idx=0)
        traj.interpolate_waypoints()
        fresh = ns.online_learner.Learner(env, traj, cost)
        syn_p = np.zeros_like(syn_cv)
        for k in range(syn_cv.shape[0]):
            getattr(fresh, case["ol_alg"])(syn_cv[k])
            syn_p[k] = fresh.p
        path = os.path.join(out_dir, "learner_%s.npz" % name)
        np.savez_compressed(
            path, alg=np.array(case["ol_alg"]), use_standoff=np.array(int(case["use_standoff"])),
            scene_args=np.array(repr(SCENE_ARGS)), sdf_checksum=np.float64(sc["sdf_grids"].astype(np.float64).sum()),
            body_points=robot.body_points, start=start, goals=goals, reach=reach, cost_vectors=cvs, p=ps,
            selected=sel, history=hist, synthetic_cv=syn_cv, synthetic_p=syn_p)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB; selected goals per trajectory:",
              [sorted(set(sel[b].tolist())) for b in range(N_TRAJ)], "synthetic leaders", sorted(set(syn_p.argmax(1).tolist())))


if __name__ == "__main__":
    main()
