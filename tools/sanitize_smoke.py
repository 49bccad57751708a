"""compute-sanitizer target: a few fused iterations in both modes on a small batch (run under
`compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitize_smoke.py`)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omg_planner_b200 import scene as S  # noqa: E402
from omg_planner_b200.config import ChompConfig  # noqa: E402
from omg_planner_b200.engine import ChompEngine  # noqa: E402
from omg_planner_b200.robot import PandaConstants  # noqa: E402

for mode in (dict(goal_set_proj=True, use_standoff=True, top_k_collision=1000),
             dict(goal_set_proj=True, use_standoff=True, top_k_collision=100),
             dict(goal_set_proj=False, use_standoff=True, top_k_collision=0)):
    sc = S.make_scene(num_objects=6, grid=32, seed=3)
    cfg = ChompConfig(**mode)
    robot = PandaConstants()
    eng = ChompEngine(robot=robot).load_scene(sc, cfg)
    xi, st, en, tails = S.make_trajectories(6, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=2)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x = dev(xi)
    for it in range(3):
        cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
        out = eng.step(cfg, x, dev(st), dev(en), dev(tails) if cfg.goal_set_proj else None, want_grad=True,
                       debug=(it == 2), want_row_obs=True)
    q = dev(xi[:, 0])
    eng.batch_obstacle_cost(q, -1, None, 0.1, -1)
    torch.cuda.synchronize()
    print(mode, "ok", float(out["info"][:, 2].sum()))

# ---- the kernels either side of the CHOMP loop (round 1e+): small instances of every entry point -----------------
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from omg_planner_b200 import core as C  # noqa: E402
from omg_planner_b200.ik import IkSolver, poses_to_targets  # noqa: E402
from omg_planner_b200.planner import Planner  # noqa: E402
from omg_planner_b200.sdf_tools import SignedDensityField  # noqa: E402

rng = np.random.RandomState(0)
robot = PandaConstants()
# trajectory initialisation: 2 and 5 knots, both modes
for K in (2, 5):
    wp = torch.from_numpy(rng.uniform(-2, 2, (7, K, 9))).cuda()
    C.interpolate_waypoints_device(wp, 13, "cubic"); C.interpolate_waypoints_device(wp, 13, "linear")
# SDF packing: mixed shapes, odd z extent, both layouts and dtypes; point-cloud field with a ragged tile
fields = []
for i, shp in enumerate([(9, 12, 7), (16, 8, 11), (5, 5, 5)]):
    data = rng.uniform(-0.1, 0.3, shp).astype(np.float32 if i % 2 == 0 else np.float64)
    if i == 1:
        raw = torch.from_numpy(np.ascontiguousarray(data.transpose(1, 0, 2))).cuda()
        fields.append(SignedDensityField(shp, np.zeros(3), 0.01, _raw=raw, _layout=1))
    else:
        fields.append(SignedDensityField(data, np.zeros(3), 0.01))
fields[2].resize(0.9)
for mx in ((16, 12, 11), (16, 12, 12)):
    C.pack_sdf_grids(fields, mx)
C.compute_sdf_from_points(rng.uniform(0.2, 0.5, (1500, 3)))
C.compute_sdf_from_points(np.zeros((0, 3)))
# inverse kinematics chains + hand poses
sol = IkSolver(robot.pose_0, robot.joint_lower_limit, robot.joint_upper_limit)
q = rng.uniform(sol.lo, sol.hi, (9, 7))
poses = sol.hand_poses(q)
tg = poses_to_targets(np.stack([poses, poses], axis=1))
sol.solve_chains(tg, rng.uniform(sol.lo, sol.hi, (3, 7)), want_steps=True)
# planner: fixed goal (persistent plan with history) and goal set with the device learner (MD and Exp)
sc = S.make_scene(num_objects=6, grid=32, seed=3)
goals, reach = S.make_goal_sets(5, 7, robot.joint_lower_limit, robot.joint_upper_limit, seed=2, spread=0.2)
for kw in (dict(goal_set_proj=False), dict(goal_set_proj=True, ol_alg="MD"), dict(goal_set_proj=True, ol_alg="Exp", use_standoff=False),
           dict(goal_set_proj=True, ol_alg="Proj")):
    cfg = ChompConfig(optim_steps=5, extra_smooth_steps=2, **kw)
    env = H.make_env(sc, cfg, robot)
    tgt = env.objects[env.target_idx]
    tgt.grasps, tgt.reach_grasps = goals, (reach if cfg.use_standoff else goals)
    traj = C.Trajectory(30, cfg=cfg, start=np.tile(S.START_CONF, (5, 1)), end=goals[:, 0])
    pl = Planner(env, traj)
    info = pl.plan(traj)
    torch.cuda.synchronize()
    print(kw, "plan ok", len(info), len(info[0]))
