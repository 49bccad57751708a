"""Optimizer: the reference's CHOMP optimizer surface (omg/optimizer.py:9-174).  One optimize() call =
schedule update on the host (written back into cfg like the reference) + ONE fused kernel launch that
evaluates the cost, applies the covariant / goal-set-projected update and the joint-limit projection."""
import time

import numpy as np


class Optimizer(object):
    def __init__(self, scene, cost):
        self.cfg = scene.config
        self.joint_lower_limit = scene.robot.joint_lower_limit
        self.joint_upper_limit = scene.robot.joint_upper_limit
        self.cost = cost
        self.step = 0
        self.time = 0.0
        self.time_elapsed = time.time()

    def reset(self):
        self.step = 0

    def update(self):
        """Schedules, written back into cfg (omg/optimizer.py:59-80)."""
        self.step += 1
        self.time_elapsed = time.time() - self.time
        self.time = time.time()
        c = self.cfg
        c.obstacle_weight = c.base_obstacle_weight * c.cost_schedule_decay ** self.step
        c.smoothness_weight = c.smoothness_base_weight * c.cost_schedule_boost ** self.step
        c.grasp_weight = c.base_grasp_weight * c.cost_schedule_decay ** self.step
        c.step_size = c.step_decay_rate ** self.step * c.base_step_size

    def report(self, curve, info):
        if not getattr(self.cfg, "report_cost", False):
            return []
        text = ["step %d lr %.5f collide %s" % (self.step, self.cfg.step_size, info["collide"]),
                "obs %.2f smooth %.2f total %.2f | grads obs %.2f smooth %.2f total %.2f | reach %.2f violate %s" % (
                    info["obs"], info["smooth"], info["cost"], info["weighted_obs_grad"],
                    info["weighted_smooth_grad"], info["grad"], info["reach"], info["violate_limit"])]
        for t in text:
            print(t)
        return text

    def optimize(self, traj, force_update=False, info_only=False):
        """One CHOMP iteration (omg/optimizer.py:115-135); returns the info dict (list of dicts when
        traj.data is batched [B,n,9])."""
        self.update()
        mode = 0 if info_only else (1 if force_update else 2)
        infos, new_xi, batched = self.cost.evaluate(traj, update_mode=mode)
        if not batched or getattr(self.cfg, "report_cost", False):   # (a batch keeps its dicts lazy: Cost.BatchInfos)
            for info in infos:
                info["text"] = self.report(traj.data, info)
        if mode != 0:
            # Trajectory.update mutates in place, Trajectory.set replaces (omg/core.py:43-57): the net effect of
            # optimize() is a replaced .data
            traj.set(new_xi if batched else new_xi[0])
        return infos if batched else infos[0]
