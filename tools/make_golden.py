#!/usr/bin/env python
"""Generate tests/golden/chomp_*.npz by running the UNMODIFIED reference Python
(omg.cost.Cost + omg.optimizer.Optimizer + robot_kinematics.forward_kinematics_parallel, imported from
/root/reference under the stubs of tools/ref_harness.py) on synthetic scenes.

BUILD-CONTAINER ONLY (needs /root/reference).  The committed .npz files are what travels.

What the fixtures pin: everything in the CHOMP iteration EXCEPT the arithmetic inside the CUDA operator
(layers/sdf_matching_loss_kernel.cu cannot be compiled here: Eigen is absent), which is
oracle/sdf_loss_ref.c for both the reference run and the oracle.  Scenes are regenerated in the tests
from (scene_args); `sdf_checksum` guards against generator drift.
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

MODES = {
    "fixed_topk": dict(goal_set_proj=False, use_standoff=True, top_k_collision=1000),
    "fixed_full": dict(goal_set_proj=False, use_standoff=True, top_k_collision=0),
    "goalset_standoff_topk": dict(goal_set_proj=True, use_standoff=True, top_k_collision=1000),
    "goalset_single_full": dict(goal_set_proj=True, use_standoff=False, top_k_collision=0),
    "goalset_standoff_topk200": dict(goal_set_proj=True, use_standoff=True, top_k_collision=200),
    # cfg.consider_finger = True: finger links stay in the top-k sum (cost.py:401-402), finger DOFs are updated
    # (core.py:47-48)
    "fixed_topk_finger": dict(goal_set_proj=False, use_standoff=True, top_k_collision=1000, consider_finger=True),
    "goalset_standoff_topk_finger": dict(goal_set_proj=True, use_standoff=True, top_k_collision=300,
                                         consider_finger=True),
}
SCENE_ARGS = dict(num_objects=6, grid=48, seed=11, grid_choices=[32, 40, 48])
N_TRAJ, N_WPT, N_ITER = 6, 30, 25
INFO_KEYS = ["obs", "smooth", "cost", "collide", "reach", "grad", "weighted_obs_grad", "weighted_smooth_grad"]
FLAG_KEYS = ["terminate", "violate_limit", "execute", "failure_terminate"]


def main():
    ns = H.load_reference()
    cfg = ns.cfg
    cfg.timeout = -1
    cfg.report_cost = False
    sc = S.make_scene(**SCENE_ARGS)
    robot = R.PandaRef()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for name, mode in MODES.items():
        if only and name not in only:
            continue
        cfg.consider_finger = False
        for k, v in mode.items():
            cfg[k] = v
        cfg.timesteps = N_WPT
        cfg.get_global_param(N_WPT)
        env = H.make_ref_env(ns, sc, robot.body_points)
        xi, st, en, tails = S.make_trajectories(N_TRAJ, N_WPT, robot.lower, robot.upper, seed=5)
        hist = np.zeros((N_TRAJ, N_ITER + 1, N_WPT, 9))
        infos = np.zeros((N_TRAJ, N_ITER, len(INFO_KEYS)))
        flags = np.zeros((N_TRAJ, N_ITER, len(FLAG_KEYS)), np.int8)
        grads = np.zeros((N_TRAJ, N_ITER, N_WPT, 9))
        slack = np.zeros((N_TRAJ, N_ITER))
        for b in range(N_TRAJ):
            cost = ns.cost.Cost(env)
            opt = ns.optimizer.Optimizer(env, cost)
            traj = H.RefTrajectory(ns, xi[b], st[b], en[b], goal_set=[en[b]], goal_idx=0)
            if cfg.goal_set_proj:
                env.objects[env.target_idx].reach_grasps = [tails[b]] if cfg.use_standoff else [en[b]]
                cost.target_obj = env.objects[env.target_idx]
            hist[b, 0] = traj.data
            # the oracle runs alongside only to record where the reference's own output is implementation-
            # defined (ties at the top-k threshold under numpy's unstable argsort): `tie_slack`
            ocfg = R.RefConfig(**mode)
            rows = None
            if cfg.goal_set_proj:
                rows = tails[b] if cfg.use_standoff else en[b][None]
            shadow = R.ChompRef(robot, sc, ocfg, xi[b], st[b], en[b], rows)
            for it in range(N_ITER):
                shadow.xi = traj.data.copy()
                slack[b, it] = shadow.step()["tie_slack"]
                info = opt.optimize(traj, force_update=True)
                hist[b, it + 1] = traj.data
                infos[b, it] = [float(info[k]) for k in INFO_KEYS]
                flags[b, it] = [int(bool(info[k])) for k in FLAG_KEYS]
                grads[b, it] = info["gradient"]
        path = os.path.join(out_dir, "chomp_%s.npz" % name)
        np.savez_compressed(
            path, mode=np.array([int(mode["goal_set_proj"]), int(mode["use_standoff"]), mode["top_k_collision"],
                                 int(mode.get("consider_finger", False))]),
            scene_args=np.array(repr(SCENE_ARGS)), sdf_checksum=np.float64(sc["sdf_grids"].astype(np.float64).sum()),
            body_points=robot.body_points, xi0=xi, start=st, end=en, tails=tails, history=hist, infos=infos,
            flags=flags, grads=grads, tie_slack=slack, info_keys=np.array(INFO_KEYS), flag_keys=np.array(FLAG_KEYS))
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB; collide range",
              infos[..., 3].min(), infos[..., 3].max(), "terminated", flags[..., 0].sum())


if __name__ == "__main__":
    main()
