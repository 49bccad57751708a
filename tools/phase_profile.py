#!/usr/bin/env python
"""Per-phase clock breakdown of the fused CHOMP kernel (diagnostic; run on the GPU box).
Writes gpurun_out/phase_profile.txt."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omg_planner_b200 import _lib, scene as S  # noqa: E402
from omg_planner_b200.config import ChompConfig  # noqa: E402
from omg_planner_b200.engine import ChompEngine, _dp  # noqa: E402
from omg_planner_b200.robot import PandaConstants  # noqa: E402

NAMES = ["stage", "fk", "cull", "compact", "points", "reduce+select", "cost", "winners", "assemble", "update",
         "limits"]


def main():
    mode = dict(goal_set_proj=True, use_standoff=True, top_k_collision=0 if "fullsum" in sys.argv else 1000)
    B = int(os.environ.get("B", 1024))
    W, O, GRID = int(os.environ.get("W", 30)), int(os.environ.get("O", 10)), int(os.environ.get("GRID", 128))
    sc = S.make_scene(num_objects=O, grid=GRID, seed=0, device="cuda" if GRID > 128 else None)
    cfg = ChompConfig(timesteps=W, **mode)
    robot = PandaConstants()
    eng = ChompEngine(robot=robot).load_scene(sc, cfg)
    xi, st, en, tails = S.make_trajectories(B, W, robot.joint_lower_limit, robot.joint_upper_limit, seed=0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x, s, e, t = dev(xi), dev(st), dev(en), dev(tails)
    prof = torch.zeros((B, 16), dtype=torch.int64, device="cuda")
    for it in range(8):
        cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
        if it == 7:
            _lib.check(eng.L.omgb_scene_set_profile(eng._h, _dp(prof)))
        out = eng.step(cfg, x, s, e, t)
    torch.cuda.synchronize()
    raw = prof.cpu().numpy()
    p = raw.astype(np.float64)
    exact = p[:, 12].copy()
    p = p[:, :12]
    d = np.diff(p, axis=1)
    info = out["info"].cpu().numpy()
    lines = ["mode %s, %d waypoints, %d objects @%d^3; per-CTA cycles (mean / median / max over %d CTAs)" % (mode, W, O, GRID, B)]
    tot = (p[:, 11] - p[:, 0])
    for k, name in enumerate(NAMES):
        lines.append("%-14s mean %9.0f  median %9.0f  max %9.0f  share %.3f" % (
            name, d[:, k].mean(), np.median(d[:, k]), d[:, k].max(), d[:, k].sum() / tot.sum()))
    lines.append("total          mean %9.0f  median %9.0f  max %9.0f" % (tot.mean(), np.median(tot), tot.max()))
    lines.append("exact operator evaluations in the points phase: mean %.1f max %.0f per trajectory" % (exact.mean(), exact.max()))
    lines.append("P_in mean %.1f  nnz mean %.1f  active link instances mean %.1f / (10 x waypoints)  limit rounds mean %.2f" % (
        info[:, 12].mean(), info[:, 13].mean(), info[:, 15].mean(), info[:, 14].mean()))
    order = np.argsort(-tot)[:6]
    lines.append("heaviest CTAs (cycles per phase):")
    for bidx in order:
        lines.append("  traj %4d total %7.0f | %s | nnz %4.0f exact %5.0f active %3.0f rounds %2.0f" % (
            bidx, tot[bidx], " ".join("%s=%d" % (nm[:4], d[bidx, k]) for k, nm in enumerate(NAMES)),
            info[bidx, 13], exact[bidx], info[bidx, 15], info[bidx, 14]))
    pct = np.percentile(tot, [50, 90, 99, 100])
    lines.append("total percentiles 50/90/99/100: %s" % np.round(pct))
    # timeline from the global timer (ns): how full the machine is over the kernel's life
    smid, t0, t1 = raw[:, 13], raw[:, 14] - raw[:, 14].min(), raw[:, 15] - raw[:, 14].min()
    span = t1.max()
    lines.append("timeline: kernel span %.1f us (first CTA start -> last CTA end); CTA duration mean %.1f us max %.1f us; "
                 "sum of CTA durations / (444 slots x span) = %.3f" % (span / 1e3, (t1 - t0).mean() / 1e3,
                                                                     (t1 - t0).max() / 1e3, (t1 - t0).sum() / (444 * span)))
    grid_t = np.linspace(0, span, 23)[1:-1]
    res = [(int(((t0 <= t) & (t1 > t)).sum())) for t in grid_t]
    lines.append("resident CTAs at %s us: %s" % ([round(t / 1e3) for t in grid_t], res))
    last_start = t0.max()
    lines.append("last CTA starts at %.1f us; CTAs starting in the first 5 us: %d; SMs used %d; per-SM busy span min/mean/max %.1f/%.1f/%.1f us" % (
        last_start / 1e3, int((t0 < 5e3).sum()), len(set(smid.tolist())),
        min(t1[smid == s_].max() for s_ in set(smid.tolist())) / 1e3,
        np.mean([t1[smid == s_].max() for s_ in set(smid.tolist())]) / 1e3,
        max(t1[smid == s_].max() for s_ in set(smid.tolist())) / 1e3))
    txt = "\n".join(lines)
    print(txt)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "phase_profile.txt"), "a") as f:
        f.write(txt + "\n\n")


if __name__ == "__main__":
    main()
