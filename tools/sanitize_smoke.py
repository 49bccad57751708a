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
