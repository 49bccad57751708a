"""TEST INFRASTRUCTURE -- CPU restatement of the outer loop of Planner.plan (omg/planner.py:600-653) on top of
oracle/chomp_ref.py and oracle/learner_ref.py, one trajectory at a time.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this.

  for t in range(optim_steps + extra_smooth_steps):
      [learner.update_goal(); selected_goals.append(goal_idx)]      goal-set mode, ol_alg not in (Baseline, Proj),
                                                                    t < optim_steps             (:614-619)
      info.append(optimize(force_update=True)); history.append(copy(traj.data))                 (:621-622)
      break if info[-1]["terminate"] and t > 0                                                  (:627-628)
  if not terminated: info.append(optimize(info_only=True))  else: del history[-1]              (:632-635)

Pinned: tests/golden/plan_*.npz hold the reference's own Planner.plan run unmodified under stubs
(tools/make_golden_plan.py); tests/test_oracle_plan.py replays them."""
import numpy as np

from . import chomp_ref as R
from . import learner_ref as LR


def plan(robot, scene, cfg, xi, start, end, goal_rows=None, goal_set=None, reach_grasps=None, goal_idx=0,
         learner=None):
    """Returns (history list, info list, selected goal list, final xi).  Fixed-goal / fixed-row mode when `learner` is
    None; otherwise the learner re-selects the goal before every one of the first optim_steps iterations."""
    opt = R.ChompRef(robot, scene, cfg, xi, start, end, goal_rows)
    history, infos, selected = [opt.xi.copy()], [], []
    for t in range(cfg.optim_steps + cfg.extra_smooth_steps):
        if learner is not None and t < cfg.optim_steps:
            learner.t += 1                                                            # online_learner.py:241
            reach = reach_grasps[:, -1, :] if cfg.use_standoff else goal_set
            cv = LR.cost_vector(robot, scene, cfg, opt.xi, goal_set, reach, learner.t)
            goal_idx = learner.update(cv, opt.xi[-1], goal_set)
            opt.end = np.array(goal_set[goal_idx]); opt.goal = opt.end.copy()        # online_learner.py:244-246
            opt.goal_rows = np.atleast_2d(reach_grasps[goal_idx] if cfg.use_standoff else goal_set[goal_idx])
            selected.append(goal_idx)
        infos.append(opt.step())
        history.append(opt.xi.copy())
        if infos[-1]["terminate"] and t > 0:
            break
    if not infos[-1]["terminate"]:
        infos.append(opt.step(info_only=True))
    else:
        del history[-1]
    return history, infos, selected, opt.xi
