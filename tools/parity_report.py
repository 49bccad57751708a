#!/usr/bin/env python
"""SURVEY 8(d) parity check at scale: from identical state, trajectories after k in {1, 10, 70} fused iterations vs the
oracle (pinned to the reference's own Python by tests/golden/), per mode: max |xi - xi_ref| over the arm DOFs, the
fraction of trajectories within 1e-4 rad, and the first diverging iteration of any failure.

  python tools/parity_report.py dump  OUT.npz     on the GPU box: runs the engine, records xi after every iteration
  python tools/parity_report.py check OUT.npz     anywhere (CPU): replays the oracle (one process per core), compares,
                                                  prints one JSON object
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# PARITY_SHAPE selects the BASELINE config whose shape is replayed: "config2" (30 waypoints, 10 objects),
# "config4" (60 waypoints, 20 objects), "config5" (50 waypoints, 30 objects); grids are kept small -- the oracle's cost
# does not depend on the grid size, the shapes of the loops do not either
_SHAPES = {"config2": (dict(num_objects=10, grid=96, seed=0), 64, 30), "config4": (dict(num_objects=20, grid=64, seed=4), 32, 60),
           "config5": (dict(num_objects=30, grid=48, seed=5), 32, 50)}
SHAPE = os.environ.get("PARITY_SHAPE", "config2")
SCENE, N_TRAJ, N_WPT = _SHAPES[SHAPE]
ITERS = 70
MODES = {
    "fixed_topk": dict(goal_set_proj=False, use_standoff=True, top_k_collision=1000),
    "fixed_full": dict(goal_set_proj=False, use_standoff=True, top_k_collision=0),
    "goalset_standoff_topk": dict(goal_set_proj=True, use_standoff=True, top_k_collision=1000),
    "goalset_single_full": dict(goal_set_proj=True, use_standoff=False, top_k_collision=0),
    "goalset_standoff_topk200": dict(goal_set_proj=True, use_standoff=True, top_k_collision=200),
    "goalset_standoff_topk_finger": dict(goal_set_proj=True, use_standoff=True, top_k_collision=1000, consider_finger=True),
}


def dump(path):
    import torch

    import helpers as H
    from omg_planner_b200 import scene as S
    from omg_planner_b200.config import ChompConfig
    from omg_planner_b200.robot import PandaConstants

    sc = S.make_scene(**SCENE)
    robot = PandaConstants()
    xi, st, en, tails = S.make_trajectories(N_TRAJ, N_WPT, robot.joint_lower_limit, robot.joint_upper_limit, seed=11)
    out = {"xi0": xi, "start": st, "end": en, "tails": tails}
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    only = os.environ.get("PARITY_MODES")
    for name, mode in MODES.items():
        if only and name not in only.split(","):
            continue
        cfg = ChompConfig(timesteps=N_WPT, **mode)
        eng = H.engine_for(sc, cfg, robot)
        rows = H.goal_rows_for(mode, tails, en)
        x = dev(xi)
        res = eng.plan(cfg, x, dev(st), dev(en), None if rows is None else dev(rows), iters=ITERS, history=True)
        h = res["hist_xi"].cpu().numpy()
        # (the big shapes keep only the checkpoints: gpurun brings back at most 64 MiB)
        out["hist_" + name] = h if (SHAPE == "config2" or os.environ.get("PARITY_FULL")) else h[[0, 9, 69]]
    np.savez_compressed(path, **out)
    print("wrote", path)


def _oracle_worker(args):
    name, mode, b, xi, st, en, rows = args
    from omg_planner_b200 import scene as S
    from oracle import chomp_ref as R

    global _SC
    try:
        sc = _SC
    except NameError:
        sc = _SC = S.make_scene(**SCENE)
    cfg = R.RefConfig(timesteps=N_WPT, **mode)
    opt = R.ChompRef(R.PandaRef(), sc, cfg, xi, st, en, rows)
    hist = np.zeros((ITERS, N_WPT, 9))
    for it in range(ITERS):
        opt.step()
        hist[it] = opt.xi
    return name, b, hist


def check(path):
    import helpers as H

    g = np.load(path)
    jobs = []
    modes = {name: mode for name, mode in MODES.items() if "hist_" + name in g.files}   # (PARITY_MODES at dump time)
    for name, mode in modes.items():
        rows = H.goal_rows_for(mode, g["tails"], g["end"])
        for b in range(N_TRAJ):
            jobs.append((name, mode, b, g["xi0"][b], g["start"][b], g["end"][b], None if rows is None else rows[b]))
    ref = {name: np.zeros((ITERS, N_TRAJ, N_WPT, 9)) for name in modes}
    with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:
        for name, b, hist in pool.imap_unordered(_oracle_worker, jobs, chunksize=4):
            ref[name][:, b] = hist
    report = {"shape": SHAPE, "scene": SCENE, "trajectories": N_TRAJ, "waypoints": N_WPT, "tolerance_rad": 1e-4, "modes": {}}
    for name, mode in modes.items():
        dofs = 9 if mode.get("consider_finger") else 7
        got = g["hist_" + name]
        full = got.shape[0] == ITERS
        err = np.abs(got - (ref[name] if full else ref[name][[0, 9, 69]]))[..., :dofs].max(axis=(2, 3))   # [iters, B]
        entry = {}
        for slot, k in enumerate((1, 10, 70)):
            e = err[k - 1] if full else err[slot]
            entry["after_%d" % k] = {"max_abs_rad": float(e.max()), "fraction_within_tolerance": float((e <= 1e-4).mean())}
        bad = np.argwhere(err > 1e-4)
        if full:
            entry["first_divergence"] = None if bad.size == 0 else {"iteration": int(bad[:, 0].min()) + 1,
                                                                    "trajectories": sorted(set(bad[:, 1].tolist()))[:8]}
        else:
            entry["trajectories_beyond_tolerance"] = sorted(set(bad[:, 1].tolist()))[:8]
        report["modes"][name] = entry
    print(json.dumps(report))


if __name__ == "__main__":
    {"dump": dump, "check": check}[sys.argv[1]](sys.argv[2])
