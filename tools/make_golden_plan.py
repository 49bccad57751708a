#!/usr/bin/env python
"""Generate tests/golden/plan_*.npz by running the UNMODIFIED reference Planner.plan (omg/planner.py:600-653, from
/root/reference under the stubs of tools/ref_harness.py) with the reference's own Cost, Optimizer and Learner on
synthetic scenes.  The Planner object is created without __init__ (which loads grasp files and runs IK); plan() itself
is the reference's.

BUILD-CONTAINER ONLY (needs /root/reference).  The committed .npz files are what travels.
"""
import builtins
import importlib
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

SCENE_ARGS = dict(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
N_WPT = 30
INFO_KEYS = ["obs", "smooth", "cost", "collide", "reach", "grad", "weighted_obs_grad", "weighted_smooth_grad"]
FLAG_KEYS = ["terminate", "violate_limit", "execute", "failure_terminate"]
CASES = {
    # fixed goal, the persistent-plan path: 50 + 20 iterations, early exit on terminate (t > 0)
    "fixed_topk": dict(goal_set_proj=False, use_standoff=True, ol_alg="MD", n_traj=5, optim_steps=50, extra=20),
    # goal set, no goal switching (ol_alg = Baseline): fixed rows, terminate needs the goal distance too
    "goalset_baseline": dict(goal_set_proj=True, use_standoff=True, ol_alg="Baseline", n_traj=3, optim_steps=50, extra=20),
    # goal set with the online learner re-selecting the goal for the first optim_steps iterations
    # (pre_terminate off: the sparse synthetic scenes let most plans stop at t = 1 otherwise)
    "goalset_md": dict(goal_set_proj=True, use_standoff=True, ol_alg="MD", n_traj=12, pick=[0, 8, 2], optim_steps=12,
                       extra=6, pre_terminate=False, spread=0.1),
    "goalset_md_terminate": dict(goal_set_proj=True, use_standoff=True, ol_alg="MD", n_traj=12, pick=[8, 1],
                                 optim_steps=12, extra=6),
    "goalset_exp_single": dict(goal_set_proj=True, use_standoff=False, ol_alg="Exp", n_traj=12, pick=[8, 5],
                               optim_steps=10, extra=4, pre_terminate=False, spread=0.1),
}
N_GOALS = 7


def main():
    ns = H.load_reference()
    sys.modules.setdefault("ycb_render.ycb_renderer", H._Anything("ycb_render.ycb_renderer"))
    planner_mod = importlib.import_module("omg.planner")
    cfg = ns.cfg
    cfg.timeout = -1
    cfg.report_cost = False
    cfg.report_time = False
    cfg.silent = True
    sc = S.make_scene(**SCENE_ARGS)
    robot = R.PandaRef()
    out_dir = os.path.join(ROOT, "tests", "golden")
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        cfg.goal_set_proj = case["goal_set_proj"]
        cfg.use_standoff = case["use_standoff"]
        cfg.ol_alg = case["ol_alg"]
        cfg.top_k_collision = 1000
        cfg.consider_finger = False
        cfg.pre_terminate = case.get("pre_terminate", True)
        cfg.optim_steps, cfg.extra_smooth_steps = case["optim_steps"], case["extra"]
        cfg.timesteps = N_WPT
        cfg.get_global_param(N_WPT)
        B, T = case["n_traj"], case["optim_steps"] + case["extra"]
        if os.environ.get("SCAN"):
            B = int(os.environ["SCAN"]); case = dict(case); case.pop("pick", None)
        env = H.make_ref_env(ns, sc, robot.body_points)
        xi, st, en, tails = S.make_trajectories(B, N_WPT, robot.lower, robot.upper, seed=5)
        goals, reach = S.make_goal_sets(B, N_GOALS, robot.lower, robot.upper, seed=case.get("goal_seed", 3),
                                        spread=case.get("spread", 0.6))
        if "pick" in case:   # keep the trajectories whose plans are interesting (chosen by scanning seeds once)
            pk = case["pick"]
            goals, reach, xi, st, en, tails = goals[pk], reach[pk], xi[pk], st[pk], en[pk], tails[pk]
            B = len(pk)
        if not cfg.goal_set_proj:
            # trajectory 1: a short move in free space -> terminates at t = 1 (the earliest the loop allows)
            en[1] = st[1] + np.array([0.05, 0.05, -0.05, 0.05, 0.0, 0.05, 0.0, 0.0, 0.0])
            xi[1] = S.clamped_cubic(st[1], en[1], N_WPT)
        hist = np.zeros((B, T + 1, N_WPT, 9)); hist_len = np.zeros(B, np.int64)
        infos = np.zeros((B, T + 1, len(INFO_KEYS))); flags = np.zeros((B, T + 1, len(FLAG_KEYS)), np.int8)
        info_len = np.zeros(B, np.int64)
        sel = -np.ones((B, T), np.int64); sel_len = np.zeros(B, np.int64)
        final = np.zeros((B, N_WPT, 9)); slack = np.zeros((B, T + 1))
        for b in range(B):
            target = env.objects[env.target_idx]
            cost = ns.cost.Cost(env)
            optim = ns.optimizer.Optimizer(env, cost)
            learner_on = cfg.goal_set_proj and cfg.ol_alg not in ("Baseline", "Proj")
            if learner_on:
                target.reach_grasps = reach[b] if cfg.use_standoff else goals[b]
                traj = H.RefTrajectory(ns, np.zeros((N_WPT, 9)), st[b], goals[b, 0], goal_set=list(goals[b]), goal_idx=0)
                traj.interpolate_waypoints()
            else:
                traj = H.RefTrajectory(ns, xi[b], st[b], en[b], goal_set=[en[b]], goal_idx=0)
                if cfg.goal_set_proj:
                    target.reach_grasps = [tails[b]]
            cost.target_obj = target
            p = planner_mod.Planner.__new__(planner_mod.Planner)
            p.cfg, p.env, p.traj, p.cost, p.optim = cfg, env, traj, cost, optim
            if learner_on:
                p.learner = ns.online_learner.Learner(env, traj, cost)   # picks the initial goal, re-initialises traj
            xi0 = traj.data.copy()
            _print = builtins.print
            builtins.print = lambda *a, **k: None
            try:
                info = p.plan(traj)
            finally:
                builtins.print = _print
            h = p.history_trajectories
            hist_len[b] = len(h); hist[b, :len(h)] = np.stack(h)
            info_len[b] = len(info)
            for k, i in enumerate(info):
                infos[b, k] = [float(i[key]) for key in INFO_KEYS]
                flags[b, k] = [int(bool(i[key])) for key in FLAG_KEYS]
            sel_len[b] = len(p.selected_goals); sel[b, :sel_len[b]] = p.selected_goals
            final[b] = traj.data
            assert np.abs(hist[b, 0] - xi0).max() == 0
            # tie slack per iteration from the oracle shadowing the recorded states
            for k in range(min(len(h) - 1, len(info))):
                rows = None
                if cfg.goal_set_proj:
                    g = int(sel[b, k]) if learner_on and k < sel_len[b] else (int(sel[b, sel_len[b] - 1]) if learner_on else 0)
                    end_k = goals[b, g] if learner_on else en[b]
                    rows = (reach[b, g] if learner_on else tails[b]) if cfg.use_standoff else end_k[None]
                else:
                    end_k = en[b]
                ocfg = R.RefConfig(goal_set_proj=cfg.goal_set_proj, use_standoff=cfg.use_standoff, top_k_collision=1000)
                shadow = R.ChompRef(robot, sc, ocfg, h[k], st[b], end_k, rows)
                shadow.iteration = k
                slack[b, k] = shadow.step()["tie_slack"]
        path = os.path.join(out_dir, "plan_%s.npz" % name)
        np.savez_compressed(
            path, pre_terminate=int(cfg.pre_terminate), goal_set_proj=int(cfg.goal_set_proj), use_standoff=int(cfg.use_standoff), ol_alg=np.array(cfg.ol_alg),
            optim_steps=cfg.optim_steps, extra_smooth_steps=cfg.extra_smooth_steps,
            scene_args=np.array(repr(SCENE_ARGS)), sdf_checksum=np.float64(sc["sdf_grids"].astype(np.float64).sum()),
            body_points=robot.body_points, xi0=hist[:, 0], start=st, end=en, tails=tails, goals=goals, reach=reach,
            history=hist, history_len=hist_len, infos=infos, flags=flags, info_len=info_len, selected=sel,
            selected_len=sel_len, final=final, tie_slack=slack, info_keys=np.array(INFO_KEYS),
            flag_keys=np.array(FLAG_KEYS))
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB; history lengths", hist_len.tolist(), "infos",
              info_len.tolist(), "terminated", [int(flags[b, info_len[b] - 1, 0]) for b in range(B)],
              "goals", [sorted(set(sel[b, :sel_len[b]].tolist())) for b in range(B)])


if __name__ == "__main__":
    main()
