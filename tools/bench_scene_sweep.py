#!/usr/bin/env python
"""BASELINE config 3 stand-in: a sweep over scenes, 256 trajectories per scene (goal-set projection with standoff,
reference defaults), every scene its own engine (its own SDF tensor, lower-bound grid and object records).  256
trajectories fill only 256 of the 444 resident CTA slots of a B200, so scenes are planned concurrently on separate CUDA
streams (the library takes the caller's stream; one persistent plan launch per scene).  One JSON line:
sequential vs concurrent whole-plan throughput."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from omg_planner_b200 import scene as S
    from omg_planner_b200.config import ChompConfig
    from omg_planner_b200.engine import ChompEngine
    from omg_planner_b200.robot import PandaConstants

    n_scenes, B, n, iters = int(os.environ.get("SCENES", 12)), 256, 30, 70
    robot = PandaConstants()
    cfg = ChompConfig(goal_set_proj=True, use_standoff=True, top_k_collision=1000)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    engines, data = [], []
    for s in range(n_scenes):
        sc = S.make_scene(num_objects=int(5 + s % 6), grid=128, seed=100 + s)      # 5-10 objects per scene
        engines.append(ChompEngine(robot=robot).load_scene(sc, cfg))
        xi, st, en, tails = S.make_trajectories(B, n, robot.joint_lower_limit, robot.joint_upper_limit, seed=s)
        data.append((xi, dev(st), dev(en), dev(tails)))
    torch.cuda.synchronize()

    def run(num_streams):
        streams = [torch.cuda.Stream() for _ in range(num_streams)]
        xs = [dev(d[0]) for d in data]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(n_scenes):
            with torch.cuda.stream(streams[s % num_streams]):
                engines[s].plan(cfg, xs[s], data[s][1], data[s][2], data[s][3], iters=iters)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, xs

    run(1)   # warm-up
    t_seq, x_seq = run(1)
    t_con, x_con = run(4)
    same = all(torch.equal(a, b) for a, b in zip(x_seq, x_con))
    total = n_scenes * B * iters
    print(json.dumps({
        "workload": "%d scenes x %d trajectories x %d waypoints, 5-10 SDFs @128^3 per scene, %d iterations, one "
                    "persistent plan launch per scene" % (n_scenes, B, n, iters),
        "sequential": {"wall_s": t_seq, "trajectory_iterations_per_s": total / t_seq},
        "four_streams": {"wall_s": t_con, "trajectory_iterations_per_s": total / t_con},
        "speedup": t_seq / t_con, "results_identical": bool(same)}))


if __name__ == "__main__":
    main()
