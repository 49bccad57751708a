"""compute-sanitizer target for the data-dependent branches of the fused step that a tiny scene does not reach: the
config-2 scene (10 SDFs @128^3), 256 trajectories, 320-thread CTAs (OMGB_STEP_CONFIG=0) -- some trajectories have more
than k non-zero points (radix select, member cost from registers) and more than 80 winners (segmented winners pass).
Run as `OMGB_STEP_CONFIG=0 compute-sanitizer --tool racecheck|memcheck python tools/sanitize_step_heavy.py`."""
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

B = int(os.environ.get("B", 256))
mode = dict(goal_set_proj=True, use_standoff=True, top_k_collision=int(os.environ.get("TOPK", 1000)))
sc = S.make_scene(num_objects=10, grid=128, seed=0)
cfg = ChompConfig(timesteps=30, **mode)
robot = PandaConstants()
eng = ChompEngine(robot=robot).load_scene(sc, cfg)
xi, st, en, tails = S.make_trajectories(1024, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a[:B])).cuda()
x, s, e, t = dev(xi), dev(st), dev(en), dev(tails)
for it in range(int(os.environ.get("ITERS", 2))):
    cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
    out = eng.step(cfg, x, s, e, t)
torch.cuda.synchronize()
info = out["info"].cpu().numpy()
print("trajectories", B, "nnz max", int(info[:, 13].max()), "nnz > k:", int((info[:, 13] > cfg.top_k_collision).sum()),
      "cost sum %.6f" % float(info[:, 2].sum()))
