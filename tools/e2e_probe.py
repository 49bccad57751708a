#!/usr/bin/env python
"""Where does the end-to-end time of omgb_chomp_step_host go?  (diagnostic; run on the GPU box)"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omg_planner_b200 import _lib, scene as S  # noqa: E402
from omg_planner_b200.config import ChompConfig  # noqa: E402
from omg_planner_b200.engine import ChompEngine, _dp, _hp, _stream  # noqa: E402
from omg_planner_b200.robot import PandaConstants  # noqa: E402


def main():
    B = 1024
    mode = dict(goal_set_proj=True, use_standoff=True, top_k_collision=1000)
    sc = S.make_scene(num_objects=10, grid=128, seed=0)
    cfg = ChompConfig(**mode)
    robot = PandaConstants()
    eng = ChompEngine(robot=robot).load_scene(sc, cfg)
    eng.set_metric(cfg)
    L = eng.L
    xi, st, en, tails = S.make_trajectories(B, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=0)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h = [pin(xi), pin(st), pin(en), pin(tails), torch.empty((B, 16), dtype=torch.float64).pin_memory()]
    d = [t.cuda() for t in h]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    prm = eng.params_from(cfg, True)
    steps = 40

    def timed(fn, label, do_flush=True):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        wall = 0.0
        for k in range(steps):
            if do_flush:
                flush.zero_()
                torch.cuda.synchronize()
            evs[k][0].record()
            t0 = time.perf_counter()
            fn()
            wall += time.perf_counter() - t0
            evs[k][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        print("%-60s events %.4f ms   host wall in call %.4f ms" % (label, ms, wall / steps * 1e3))

    s = _stream()
    dev_args = [_dp(t) for t in d]
    host_args = [ctypes.c_void_p(t.data_ptr()) for t in h]

    def k_dev():
        L.omgb_chomp_step(eng._h, ctypes.byref(prm), B, dev_args[0], dev_args[1], dev_args[2], dev_args[3], None, None,
                          dev_args[4], None, None, None, s)

    def k_map():
        L.omgb_chomp_step(eng._h, ctypes.byref(prm), B, host_args[0], host_args[1], host_args[2], host_args[3], None, None,
                          host_args[4], None, None, None, s)

    def k_map_xi_dev_rest():
        L.omgb_chomp_step(eng._h, ctypes.byref(prm), B, host_args[0], dev_args[1], dev_args[2], dev_args[3], None, None,
                          dev_args[4], None, None, None, s)

    def host_call():
        L.omgb_chomp_step_host(eng._h, ctypes.byref(prm), B, host_args[0], host_args[1], host_args[2], host_args[3],
                               host_args[4], s)

    def eng_call():
        eng.step_host(cfg, h[0].numpy(), h[1].numpy(), h[2].numpy(), h[3].numpy())

    def combo(mask):
        ar = [host_args[i] if (mask >> i) & 1 else dev_args[i] for i in range(5)]

        def f():
            L.omgb_chomp_step(eng._h, ctypes.byref(prm), B, ar[0], ar[1], ar[2], ar[3], None, None, ar[4], None, None,
                              None, s)
        return f

    if "combos" in sys.argv:
        names = ["xi", "start", "end", "goal", "info"]
        for mask in (0, 1, 2, 4, 8, 16, 1 | 16, 1 | 2 | 4, 1 | 8, 31):
            timed(combo(mask), "kernel, mapped: " + ",".join(n for i, n in enumerate(names) if (mask >> i) & 1))
        return
    timed(k_dev, "kernel, device buffers (async launch)")
    timed(k_map, "kernel, all buffers mapped pinned host (async launch)")
    timed(k_map_xi_dev_rest, "kernel, xi mapped host, rest device (async launch)")
    for m, nm in ((0, "zero-copy"), (1, "staged"), (2, "staged pipelined")):
        eng.set_host_mode(m)
        timed(host_call, "omgb_chomp_step_host raw ctypes, %s" % nm)
        timed(eng_call, "ChompEngine.step_host, %s" % nm)
    eng.set_host_mode(0)
    timed(host_call, "omgb_chomp_step_host raw ctypes zero-copy, no L2 flush", do_flush=False)


if __name__ == "__main__":
    main()
