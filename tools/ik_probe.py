"""GPU probe: omgb_ik_solve vs the C restatement / the reference's KDL on random reachable and unreachable poses."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omg_planner_b200.ik import IkSolver, poses_to_targets
from omg_planner_b200.robot import PandaConstants
from oracle import kdl_ik_ref as K
import torch

r = PandaConstants()
ch = K.PandaChain(r.pose_0, r.joint_lower_limit, r.joint_upper_limit)
sol = IkSolver(r.pose_0, r.joint_lower_limit, r.joint_upper_limit)
rng = np.random.RandomState(3)
P, S = int(os.environ.get("P", 400)), 13
poses = np.stack([ch.fk_hand(rng.uniform(ch.lo, ch.hi)) for _ in range(P)])
poses[::7, :3, 3] += rng.uniform(-0.4, 0.4, (len(poses[::7]), 3))
np.testing.assert_allclose(sol.hand_poses(np.stack([rng.uniform(ch.lo, ch.hi) for _ in range(5)])).shape, (5, 4, 4))
seeds = np.concatenate([[np.array([0.0, -1.285, 0, -2.356, 0.0, 1.571, 0.785])], rng.uniform(ch.lo, ch.hi, (S - 1, 7))])
tg = poses_to_targets(poses)[:, None]
sols, solved, steps = sol.solve_chains(tg, seeds, want_steps=True)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(3):
    sol.solve_chains(tg, seeds)
torch.cuda.synchronize(); dt = (time.time() - t0) / 3
agree = mism = both = 0; worst = 0.0; diffs = []
t1 = time.time()
for p in range(P):
    for s in range(S):
        q, rc, its, raw = ch.ik(tg[p, 0, :3], tg[p, 0, 3:], seeds[s])
        ok_g = solved[p, s] == 1
        if ok_g == (rc >= 0):
            agree += 1
            if ok_g:
                both += 1
                d = np.abs(sols[p, s, 0] - q).max(); diffs.append(d); worst = max(worst, d)
        else:
            mism += 1
cpu = time.time() - t1
diffs = np.array(diffs)
print("problems %d  status agree %d  mismatch %d  both-solved %d" % (P * S, agree, mism, both))
print("solution |dq|: max %.3e  p99 %.3e  median %.3e  bit-identical %d" % (worst, np.percentile(diffs, 99), np.median(diffs), (diffs == 0).sum()))
print("GPU %.2f ms for %d chains (incl. H2D/D2H); oracle C %.1f ms (1 core) -> %.0fx" % (dt * 1e3, P * S, cpu * 1e3, cpu / dt))
print("steps histogram (gpu):", np.bincount(np.minimum(steps.reshape(-1), 100) // 10))
