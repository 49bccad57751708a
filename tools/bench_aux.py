#!/usr/bin/env python
"""Timings of the kernels either side of the CHOMP loop: SDF packing (HBM-bound: the one kernel of
this repo whose DRAM bytes equal its algorithmic bytes), point-cloud distance field, trajectory initialisation, batched
inverse kinematics.  CUDA events on the launch stream, 3 warm-ups, L2 flushed between timed launches.
`python tools/bench_aux.py` prints one JSON object; bench.py embeds the same object as "aux_kernels" at N=1."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time(fn, flush, reps=5, warm=3):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def run_aux(hbm_peak_gbs=None, ik_oracle_sample=60):
    import ctypes

    import torch

    from omg_planner_b200 import _lib
    from omg_planner_b200 import core as C
    from omg_planner_b200.ik import IkSolver, poses_to_targets
    from omg_planner_b200.robot import PandaConstants
    from omg_planner_b200.sdf_tools import SignedDensityField

    L = _lib.lib()
    vp = ctypes.c_void_p
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    out = {}
    rng = np.random.RandomState(0)

    # ---- omgb_sdf_pack: config-4 scene, 20 objects stored the .pth way at mixed sizes up to 256^3 ----------------
    fields, raw_bytes = [], 0
    for i in range(20):
        shp = (256, 256, 256) if i == 0 else tuple(int(v) for v in rng.randint(160, 257, 3))
        raw = torch.rand((shp[1], shp[0], shp[2]), dtype=torch.float32, device="cuda")
        fields.append(SignedDensityField(shp, np.zeros(3), 0.004, _raw=raw, _layout=1))
        raw_bytes += raw.numel() * 4
    mx = np.array([f.shape for f in fields]).max(0)
    dst_bytes = 20 * int(mx[0]) * int(mx[1]) * int(mx[2]) * 4
    ms = _time(lambda: C.pack_sdf_grids(fields, mx), flush)
    # (pack_sdf_grids allocates its output; time the launch alone too)
    dst = torch.empty((20, int(mx[0]), int(mx[1]), int(mx[2])), dtype=torch.float32, device="cuda")
    table = (_lib.SdfSource * 20)()
    for i, f in enumerate(fields):
        table[i].data = f.raw.data_ptr()
        table[i].shape[0], table[i].shape[1], table[i].shape[2] = f.nx, f.ny, f.nz
        table[i].layout, table[i].dtype, table[i].scale = 1, 0, 1.0
    st = vp(torch.cuda.current_stream().cuda_stream)
    ms_k = _time(lambda: L.omgb_sdf_pack(table, 20, int(mx[0]), int(mx[1]), int(mx[2]), vp(dst.data_ptr()), st), flush)
    gbs = (raw_bytes + dst_bytes) / (ms_k * 1e-3) / 1e9
    out["sdf_pack"] = {"workload": "20 objects, .pth layout, up to 256^3 -> [20,%d,%d,%d] fp32" % tuple(mx),
                       "ms": ms_k, "ms_with_allocation": ms, "algorithmic_bytes": raw_bytes + dst_bytes,
                       "achieved_gbs": gbs, "bound": "hbm",
                       "frac_of_hbm_peak": (gbs / hbm_peak_gbs) if hbm_peak_gbs else None}
    del dst, fields
    torch.cuda.empty_cache()

    # ---- omgb_sdf_loss: the drop-in operator (omg_cuda.sdf_loss_forward) on config 2's point count --------------------
    from omg_planner_b200 import scene as S
    from omg_planner_b200.cost import se3_inverse_f32
    from omg_planner_b200.engine import sdf_loss_forward

    sc = S.make_scene(num_objects=10, grid=128, seed=0)
    N = 1024 * 30 * 150
    pts_op = torch.from_numpy(rng.uniform([0.2, -0.5, -0.1], [0.9, 0.5, 0.6], (N, 3)).astype(np.float32)).cuda()
    args = [torch.from_numpy(np.stack([se3_inverse_f32(m) for m in sc["pose_mats"]])).cuda(),
            torch.from_numpy(sc["sdf_grids"]).cuda(), torch.from_numpy(sc["sdf_limits"]).cuda(), pts_op,
            torch.full((10,), 0.2).cuda(), torch.ones(10).cuda(), torch.full((10,), 0.01).cuda(), torch.zeros(10).cuda()]
    ms = _time(lambda: sdf_loss_forward(*args), flush, reps=3, warm=1)
    pot, grad, col = sdf_loss_forward(*args)
    out["sdf_loss_operator"] = {"workload": "%d points (1024 x 30 x 150) x 10 objects @128^3, uniform in the workspace box" % N,
                                "ms": ms, "points_per_s": N / (ms * 1e-3), "point_object_pairs_per_s": 10 * N / (ms * 1e-3),
                                "nonzero_potentials": int((pot > 0).sum().item())}
    del args, pts_op, pot, grad, col
    torch.cuda.empty_cache()

    # ---- omgb_point_sdf: table-top cloud -----------------------------------------------------------------------
    pts = rng.uniform([0.2, -0.5, 0.0], [1.0, 0.5, 0.6], (20000, 3))
    f = C.compute_sdf_from_points(pts)
    ms = _time(lambda: C.compute_sdf_from_points(pts), flush, reps=3, warm=1)
    pairs = f.nx * f.ny * f.nz * pts.shape[0]
    out["point_sdf"] = {"workload": "%d points, grid %dx%dx%d (2 cm)" % (pts.shape[0], f.nx, f.ny, f.nz), "ms": ms,
                        "voxel_point_pairs_per_s": pairs / (ms * 1e-3), "fp64_ops_per_s": 9 * pairs / (ms * 1e-3),
                        "bound": "fp64 issue"}

    # ---- omgb_traj_interpolate ---------------------------------------------------------------------------------
    wp = torch.from_numpy(rng.uniform(-2, 2, (8192, 2, 9))).cuda()
    ms = _time(lambda: C.interpolate_waypoints_device(wp, 60), flush)
    out["traj_interpolate"] = {"workload": "8192 trajectories x 60 waypoints", "ms": ms,
                               "trajectories_per_s": 8192 / (ms * 1e-3)}

    # ---- omgb_ik_solve: 300 grasp poses x 13 seeds x (1 + 5) chained solves ------------------------------------------
    robot = PandaConstants()
    sol = IkSolver(robot.pose_0, robot.joint_lower_limit, robot.joint_upper_limit)
    lo, hi = sol.lo, sol.hi
    q = rng.uniform(lo, hi, (300, 7))
    base = sol.hand_poses(q)
    back = np.tile(np.eye(4), (6, 1, 1))
    back[:, 2, 3] = -0.08 * np.array([4, 0, 1, 2, 3, 4]) / 5.0
    targets = poses_to_targets(np.matmul(base[:, None], back[None]))
    seeds = np.concatenate([[np.array([0.0, -1.285, 0, -2.356, 0.0, 1.571, 0.785])], rng.uniform(lo, hi, (12, 7))])
    d_t, d_s = torch.from_numpy(targets).cuda(), torch.from_numpy(seeds).cuda()
    sols = torch.zeros((300, 13, 6, 7), dtype=torch.float64, device="cuda")
    solved = torch.zeros((300, 13), dtype=torch.int32, device="cuda")
    ms = _time(lambda: L.omgb_ik_solve(sol.frames.ctypes.data, lo.ctypes.data, hi.ctypes.data, vp(d_t.data_ptr()), 300,
                                       6, vp(d_s.data_ptr()), 13, vp(sols.data_ptr()), vp(solved.data_ptr()), None, st),
               flush, reps=3, warm=1)
    ik = {"workload": "300 grasp poses x 13 seeds, chains of 6 solves (standoff pattern)", "ms": ms,
          "chains_per_s": 3900 / (ms * 1e-3), "fully_solved": int((solved == 6).sum().item())}
    try:   # the CPU checker on a bounded sample, one core
        from oracle import kdl_ik_ref as K

        ch = K.PandaChain(robot.pose_0, robot.joint_lower_limit, robot.joint_upper_limit)
        t0 = time.perf_counter()
        cnt = 0
        for p in range(ik_oracle_sample):
            ch.ik_chain(targets[p], seeds[p % 13])
            cnt += 1
        dt = time.perf_counter() - t0
        ik["cpu_restatement_chains_per_s_one_core"] = cnt / dt
        if K.have_ref():
            t0 = time.perf_counter()
            for p in range(ik_oracle_sample):
                qq = seeds[p % 13]
                for t in range(6):
                    r, rc, raw = ch.ref_ik(targets[p, t, :3], targets[p, t, 3:], qq)
                    if rc < 0:
                        break
                    qq = r
            ik["reference_kdl_chains_per_s_one_core"] = ik_oracle_sample / (time.perf_counter() - t0)
    except Exception as e:   # noqa: BLE001
        ik["cpu_note"] = repr(e)
    out["ik_solve"] = ik
    # the same with the grasp set up-sampled 10x (cfg.y_upsample): throughput when there are enough chains to fill
    # the machine (3000 poses x 13 seeds = 39000 chains)
    t10 = np.tile(targets, (10, 1, 1)) + rng.normal(0, 1e-3, (3000, 1, 7)) * np.array([1, 1, 1, 0, 0, 0, 0])
    d_t10 = torch.from_numpy(np.ascontiguousarray(t10)).cuda()
    sols10 = torch.zeros((3000, 13, 6, 7), dtype=torch.float64, device="cuda")
    solved10 = torch.zeros((3000, 13), dtype=torch.int32, device="cuda")
    ms = _time(lambda: L.omgb_ik_solve(sol.frames.ctypes.data, lo.ctypes.data, hi.ctypes.data, vp(d_t10.data_ptr()),
                                       3000, 6, vp(d_s.data_ptr()), 13, vp(sols10.data_ptr()), vp(solved10.data_ptr()),
                                       None, st), flush, reps=2, warm=1)
    out["ik_solve_upsampled"] = {"workload": "3000 grasp poses x 13 seeds, chains of 6 solves", "ms": ms,
                                 "chains_per_s": 39000 / (ms * 1e-3), "fully_solved": int((solved10 == 6).sum().item())}
    return out


if __name__ == "__main__":
    peak = None
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    print(json.dumps(run_aux(peak)))
