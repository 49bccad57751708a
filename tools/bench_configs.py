#!/usr/bin/env python
"""BASELINE.json configs 3, 4 and 5 at their full sizes, as blocks of bench.py's JSON line (the headline `value`
stays config 2).  Every block carries value / ms_per_step / its own roofline / the timed all-gather (N > 1) and
`parity_frac_within_1e-4`: the fraction of an oracle-checked subsample of the block's own trajectories whose joint
angles agree with the CPU oracle within 1e-4 rad per waypoint -- the oracle is the checker here, never
the thing measured.

  config4  synthetic 8192 trajectories x 60 waypoints, 20 SDFs @256^3 (1.34 GB, replicated), trajectories split over
           WORLD_SIZE (strong scaling: 1024 per GPU at N = 8), reference default mode, one all-gather of final costs
  config5  kitchen-like scene: 512 trajectories x 50 waypoints (bullet/panda_scene.py:572-573), 30 obstacle SDFs,
           online goal re-weighting (MD learner, omg/config.py:67) -> whole Planner.plan, split over WORLD_SIZE
  config3  the -exp sweep (omg/core.py:860-885: use_standoff False): 100 scenes x 256 trajectories x 30 waypoints,
           goal set on, scenes split over WORLD_SIZE

Also runnable on its own:  python tools/bench_configs.py [config3|config4|config5 ...]"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DEFAULT_MODE = dict(goal_set_proj=True, use_standoff=True, top_k_collision=1000)
METRIC_UNIT = "trajectory-iterations/s"


# ----------------------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def _max_over_ranks(ms):
    import torch
    d = _dist()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if d is not None:
        d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def _barrier():
    import torch
    torch.cuda.synchronize()
    d = _dist()
    if d is not None:
        d.barrier()
    torch.cuda.synchronize()


def _shard(total, rank, world):
    from omg_planner_b200.dist import shard_range
    lo, hi = shard_range(total, rank, world)
    return lo, hi


def _host_scene(sc):
    """numpy copy of a device-generated scene dict (for the oracle)."""
    import torch
    out = dict(sc)
    if torch.is_tensor(sc["sdf_grids"]):
        out["sdf_grids"] = sc["sdf_grids"].cpu().numpy()
    return out


_PAR = {}


def _pool_init(d):
    """spawned worker: arguments from a pickle, the (possibly 1.34 GB) SDF tensor memory-mapped from a .npy file"""
    import pickle
    with open(os.path.join(d, "args.pkl"), "rb") as f:
        sc, rest = pickle.load(f)
    sc["sdf_grids"] = np.load(os.path.join(d, "grid.npy"), mmap_mode="r")
    _PAR["args"] = (sc,) + tuple(rest)


def _pool_map(fn, count, args):
    """Run the CPU oracle for `count` trajectories on the host cores.  The workers are SPAWNED (this process holds a
    CUDA context and NCCL threads; they only ever run numpy + the C operator) and get args[0], the scene, through a
    temporary directory: the big grid as a memory-mapped file shared by all of them."""
    import pickle
    import shutil
    import tempfile

    workers = max(1, min(count, (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 4)))
    d = tempfile.mkdtemp(prefix="omgb_oracle_")
    try:
        sc = dict(args[0])
        np.save(os.path.join(d, "grid.npy"), np.ascontiguousarray(sc.pop("sdf_grids"), dtype=np.float32))
        with open(os.path.join(d, "args.pkl"), "wb") as f:
            pickle.dump((sc, tuple(args[1:])), f)
        with mp.get_context("spawn").Pool(workers, initializer=_pool_init, initargs=(d,)) as pool:
            res = pool.map(fn, range(count))
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return res, workers


def _oracle_steps_worker(b):
    from oracle import chomp_ref as R
    sc, mode, n, xi, st, en, rows, checkpoints = _PAR["args"]
    cfg = R.RefConfig(timesteps=n, **mode)
    opt = R.ChompRef(R.PandaRef(), sc, cfg, xi[b], st[b], en[b], None if rows is None else rows[b])
    out, p_in = [], []
    for it in range(max(checkpoints)):
        info = opt.step()
        p_in.append(info.get("p_in", -1))
        if it + 1 in checkpoints:
            out.append(opt.xi.copy())
    return b, np.stack(out), p_in


def _oracle_trace_worker(b):
    """xi after EVERY iteration (tests/test_gpu_configs_fullsize.py)."""
    from oracle import chomp_ref as R
    sc, mode, n, xi, st, en, rows, iters = _PAR["args"]
    opt = R.ChompRef(R.PandaRef(), sc, R.RefConfig(timesteps=n, **mode), xi[b], st[b], en[b], rows[b])
    out, pin = [], []
    for _ in range(iters):
        info = opt.step()
        out.append(opt.xi.copy())
        pin.append(info["p_in"])
    return b, np.stack(out), pin


def _oracle_parity_steps(sc_host, mode, n, xi, st, en, rows, dev_states, checkpoints=(1, 10)):
    """dev_states[k]: device xi [S,n,9] after checkpoints[k] iterations from the same state."""
    S = xi.shape[0]
    t0 = time.perf_counter()
    out, workers = _pool_map(_oracle_steps_worker, S, (sc_host, mode, n, xi, st, en, rows, tuple(checkpoints)))
    res = dict((b, h) for b, h, _ in out)
    err = np.zeros((len(checkpoints), S))
    for b in range(S):
        for k in range(len(checkpoints)):
            err[k, b] = np.abs(dev_states[k][b] - res[b][k])[:, :7].max()
    return {"sample_trajectories": S, "iterations_checked": list(checkpoints),
            "max_abs_rad": [float(e.max()) for e in err],
            "parity_frac_within_1e-4": float((err.max(0) <= 1e-4).mean()),
            "outliers": [int(b) for b in np.nonzero(err.max(0) > 1e-4)[0]],
            "oracle": "oracle/chomp_ref.py + oracle/sdf_loss_ref.c on %d host processes, %.1f s" % (
                workers, time.perf_counter() - t0)}


def _roofline(p_in_per_launch, B, n, c, kern_ms, peak, peak_src, kernel, traffic_key):
    bytes_per_launch = 128.0 * p_in_per_launch + B * (8.0 * n * 9 + 4.0 * (2 + c) * 9)
    achieved = bytes_per_launch / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(traffic_key)
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "kernel": kernel,
            "algorithmic_bytes_per_launch": bytes_per_launch, "p_in_per_launch": p_in_per_launch, "kernel_ms": kern_ms}


# ----------------------------------------------------------------------------------------------------------------
# config 4
# ----------------------------------------------------------------------------------------------------------------
def run_config4(rank, world, peak, peak_src, steps=10, warmup=6, total=8192, n=60, objects=20, grid=256, parity=True,
                plan_iters=70):
    import torch

    from omg_planner_b200 import scene as S
    from omg_planner_b200.config import ChompConfig
    from omg_planner_b200.engine import ChompEngine
    from omg_planner_b200.robot import PandaConstants

    dev = torch.device("cuda", torch.cuda.current_device())
    lo, hi = _shard(total, rank, world)
    B = hi - lo
    t0 = time.perf_counter()
    sc = S.make_scene(num_objects=objects, grid=grid, seed=4, device=dev)
    t_scene = time.perf_counter() - t0
    cfg = ChompConfig(timesteps=n, **DEFAULT_MODE)
    robot = PandaConstants()
    eng = ChompEngine(robot=robot).load_scene(sc, cfg)
    c = cfg.constraint_rows
    # every rank draws the whole batch's parameters cheaply? no: only its shard (seed = global trajectory index)
    xi0, st, en, tails = S.make_trajectories(B, n, robot.joint_lower_limit, robot.joint_upper_limit, seed=40 + rank)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    xi, d_st, d_en, d_tails = to(xi0), to(st), to(en), to(tails)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    it = 0

    def one_step(ev=None):
        nonlocal it
        cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
        it += 1
        flush.zero_()
        if ev is not None:
            ev[0].record()
        out = eng.step(cfg, xi, d_st, d_en, d_tails)
        if ev is not None:
            ev[1].record()
        return out

    # warm-up = the timed loop's exact allocation pattern (two info tensors alive at a time, the P_in sums): a
    # cudaMalloc by torch's caching allocator inside a timed step would stall the device for milliseconds
    wev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(warmup, 3))]
    pins = []
    for k in range(len(wev)):
        out = one_step(wev[k])
        pins.append(out["info"][:, 12].sum())
    _barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    pins = []
    for k in range(steps):
        out = one_step(evs[k])
        pins.append(out["info"][:, 12].sum())
    _barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in evs)
    print("config4 rank %d: per-step device ms %s" % (rank, " ".join("%.3f" % s.elapsed_time(e) for s, e in evs)),
          file=sys.stderr)
    dev_ms_max = _max_over_ranks(dev_ms)
    p_in = float(torch.stack(pins).mean().item())

    # the one collective of the path: all-gather of the final per-trajectory costs, timed on the device
    allgather_ms = None
    final_cost = out["info"][:, 2].contiguous()
    d = _dist()
    if d is not None:
        from omg_planner_b200 import dist as D
        D.all_gather_costs(final_cost)            # warm-up (communicator / buffers)
        _barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gathered = D.all_gather_costs(final_cost)
        e1.record()
        torch.cuda.synchronize()
        assert gathered.shape[0] == total
        allgather_ms = _max_over_ranks(e0.elapsed_time(e1))

    # whole plan (reference schedule) as one persistent launch, L2 warm
    xp = to(xi0)
    eng.plan(cfg, xp, d_st, d_en, d_tails, iters=plan_iters)
    xp.copy_(to(xi0))
    _barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.plan(cfg, xp, d_st, d_en, d_tails, iters=plan_iters)
    e1.record()
    _barrier()
    plan_ms_max = _max_over_ranks(e0.elapsed_time(e1))

    block = {
        "workload": "config4: %d traj x %d wpt x 7-DOF Panda, %d synthetic SDFs @%d^3 (%.2f GB replicated per GPU), "
                    "goal-set+standoff top_k=1000, trajectories split over %d GPU(s) (%d per GPU)" % (
                        total, n, objects, grid, sc["sdf_grids"].numel() * 4 / 1e9, world, B),
        "value": total * steps / (dev_ms_max * 1e-3), "unit": METRIC_UNIT, "scaling": "strong",
        "ms_per_step": dev_ms_max / steps, "steps": steps, "warmup": max(warmup, 3), "batch_per_gpu": B,
        "l2": "flushed between timed steps", "allgather_ms": allgather_ms,
        "allgather": None if allgather_ms is None else "NCCL all_gather_into_tensor of %d x fp64 final costs per rank, "
                                                       "CUDA events, max over ranks" % B,
        "value_incl_allgather": None if allgather_ms is None else total * steps / ((dev_ms_max + allgather_ms) * 1e-3),
        "plan_persistent": {"value": total * plan_iters / (plan_ms_max * 1e-3), "unit": METRIC_UNIT,
                            "iterations": plan_iters, "ms_per_iteration": plan_ms_max / plan_iters},
        "roofline": _roofline(p_in, B, n, c, dev_ms / steps, peak, peak_src, "chomp_step_kernel<topk>",
                              "default:flush:%dx%dx%dx%d" % (B, n, objects, grid)),
        "scene_build_s": t_scene,
    }
    if parity and rank == 0:
        Sn = 8
        xs = to(xi0[:Sn])
        states = []
        pcfg = ChompConfig(timesteps=n, **DEFAULT_MODE)
        for k in range(10):
            pcfg.obstacle_weight, pcfg.smoothness_weight, pcfg.step_size = pcfg.schedule(k + 1)
            eng.step(pcfg, xs, d_st[:Sn].contiguous(), d_en[:Sn].contiguous(), d_tails[:Sn].contiguous())
            if k + 1 in (1, 10):
                states.append(xs.cpu().numpy())
        block["parity"] = _oracle_parity_steps(_host_scene(sc), DEFAULT_MODE, n, xi0[:Sn], st[:Sn], en[:Sn], tails[:Sn],
                                               states)
        block["parity_frac_within_1e-4"] = block["parity"]["parity_frac_within_1e-4"]
    del eng, flush
    return block


# ----------------------------------------------------------------------------------------------------------------
# goal-set plans (configs 3 and 5): the public Planner API, device-resident learner
# ----------------------------------------------------------------------------------------------------------------
def _env_for(sc, cfg, robot):
    """What Cost/Optimizer/Planner read from omg.core.Env (SURVEY 8b 'Scene inputs'); sdf_torch stays on the device."""
    import types

    import torch

    env = types.SimpleNamespace()
    env.config = cfg
    env.target_idx = sc["target_idx"]
    env.objects = [types.SimpleNamespace(name=nm, pose_mat=np.array(sc["pose_mats"][i]), attached=False, reach_grasps=[],
                                         grasps=[]) for i, nm in enumerate(sc["names"])]
    g = sc["sdf_grids"]
    env.sdf_torch = g if torch.is_tensor(g) else torch.from_numpy(g).cuda()
    env.sdf_limits = torch.from_numpy(sc["sdf_limits"]).to(env.sdf_torch.device)
    rk = types.SimpleNamespace(_pose_0=robot.pose_0, _tip2joint=robot.tip2joint, _joint_axis=robot.joint_axis,
                               _joint_origin=robot.joint_axis, center_offset=robot.center_offset)
    env.robot = types.SimpleNamespace(robot_kinematics=rk, collision_points=robot.collision_points,
                                      joint_lower_limit=robot.joint_lower_limit, joint_upper_limit=robot.joint_upper_limit)
    return env


def _goalset_planner(sc, cfg, robot, goals, reach, n):
    from omg_planner_b200 import core as C
    from omg_planner_b200 import scene as S
    from omg_planner_b200.planner import Planner

    env = _env_for(sc, cfg, robot)
    target = env.objects[env.target_idx]
    target.grasps = goals
    target.reach_grasps = reach if cfg.use_standoff else goals
    B = goals.shape[0]
    traj = C.Trajectory(n, cfg=cfg, start=np.tile(S.START_CONF, (B, 1)), end=goals[:, 0])
    return Planner(env, traj), env, traj


def _fresh_traj(planner, env, cfg, goals, n):
    from omg_planner_b200 import core as C
    from omg_planner_b200 import scene as S

    traj = C.Trajectory(n, cfg=cfg, start=np.tile(S.START_CONF, (goals.shape[0], 1)), end=goals[:, 0])
    traj.goal_set = goals
    planner.update(env, traj)
    return traj


def _oracle_plan_worker(b):
    from omg_planner_b200 import scene as S
    from oracle import chomp_ref as R
    from oracle import learner_ref as LR
    from oracle import planner_ref as P

    sc, cfg_kw, n, xi0, g0, goals, reach = _PAR["args"]
    cfg = R.RefConfig(timesteps=n, top_k_collision=1000, **cfg_kw)
    G = goals.shape[1]
    learner = LR.LearnerRef(cfg, G)
    standoff = cfg.use_standoff
    rows = reach[b][g0[b]] if standoff else goals[b][g0[b]][None]
    hist, infos, sel, final = P.plan(R.PandaRef(), sc, cfg, xi0[b], S.START_CONF, goals[b][g0[b]], rows, goal_set=goals[b],
                                     reach_grasps=reach[b] if standoff else goals[b][:, None, :], goal_idx=int(g0[b]),
                                     learner=learner)
    return b, np.stack(hist), [int(s) for s in sel]


def _oracle_parity_plan(sc_host, cfg_kw, n, xi0, g0, goals, reach, dev_hist, dev_sel):
    S = xi0.shape[0]
    t0 = time.perf_counter()
    res, workers = _pool_map(_oracle_plan_worker, S, (sc_host, cfg_kw, n, xi0, g0, goals, reach))
    err, same = np.zeros(S), 0
    for b, hist, sel in res:
        m = min(len(hist), len(dev_hist[b]))
        err[b] = np.abs(np.asarray(dev_hist[b])[:m] - hist[:m])[..., :7].max() if len(hist) == len(dev_hist[b]) else np.inf
        same += int(list(dev_sel[b]) == sel)
    return {"sample_trajectories": S, "iterations_checked": int(cfg_kw["optim_steps"] + cfg_kw["extra_smooth_steps"]),
            "max_abs_rad": float(err.max()), "parity_frac_within_1e-4": float((err <= 1e-4).mean()),
            "selected_goal_sequences_identical": "%d/%d" % (same, S),
            "oracle": "oracle/planner_ref.py + learner_ref.py + chomp_ref.py on %d host processes, %.1f s" % (
                workers, time.perf_counter() - t0)}


def run_config5(rank, world, total=512, n=50, objects=30, goals_per_traj=20, parity=True):
    import torch

    from omg_planner_b200 import _lib
    from omg_planner_b200 import scene as S
    from omg_planner_b200.config import ChompConfig
    from omg_planner_b200.robot import PandaConstants

    dev = torch.device("cuda", torch.cuda.current_device())
    lo, hi = _shard(total, rank, world)
    B = hi - lo
    sc = S.make_scene(num_objects=objects, grid=160, seed=5, grid_choices=[64, 96, 128, 160], device=dev)
    robot = PandaConstants()
    cfg = ChompConfig(timesteps=n, goal_set_proj=True, use_standoff=True, ol_alg="MD", pre_terminate=False)
    goals, reach = S.make_goal_sets(B, goals_per_traj, robot.joint_lower_limit, robot.joint_upper_limit, seed=50 + rank,
                                    spread=0.3)
    planner, env, traj = _goalset_planner(sc, cfg, robot, goals, reach, n)
    iters = cfg.optim_steps + cfg.extra_smooth_steps
    walls, devs = [], []
    launches = 0
    for rep in range(3):   # first pass = warm-up
        traj = _fresh_traj(planner, env, cfg, goals, n)
        _barrier()
        l0 = int(_lib.lib().omgb_launch_count())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        planner.plan(traj)
        e1.record()
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
        devs.append(e0.elapsed_time(e1))
        launches = int(_lib.lib().omgb_launch_count()) - l0
    wall_ms = _max_over_ranks(min(walls[1:]) * 1e3)
    block = {
        "workload": "config5: kitchen-like scene, %d traj x %d wpt, %d obstacle SDFs (64^3..160^3, padded to %s), goal sets "
                    "of %d with online re-weighting (MD learner), standoff, 50 + 20 iterations, split over %d GPU(s) "
                    "(%d per GPU)" % (total, n, objects, "x".join(str(v) for v in sc["sdf_grids"].shape[1:]),
                                      goals_per_traj, world, B),
        "value": total * iters / (wall_ms * 1e-3), "unit": METRIC_UNIT, "scaling": "strong",
        "ms_per_step": wall_ms / iters, "steps": iters, "batch_per_gpu": B,
        "timing": "host wall clock around Planner.plan (numpy trajectory in, histories / info lists / selected goals out: "
                  "H2D and D2H inside), best of 2 after a warm-up plan, max over ranks",
        "gpu_launches_per_plan": launches,
        "api": "omg_planner_b200.planner.Planner.plan: per iteration omgb_goal_costs -> omgb_learner_update -> "
               "omgb_chomp_plan_step on one stream",
    }
    if parity and rank == 0:
        Sn = 4
        pk = dict(goal_set_proj=True, use_standoff=True, ol_alg="MD", pre_terminate=False, optim_steps=10,
                  extra_smooth_steps=4)
        pcfg = ChompConfig(timesteps=n, **pk)
        p2, env2, traj2 = _goalset_planner(sc, pcfg, robot, goals[:Sn], reach[:Sn], n)
        xi0, g0 = np.array(traj2.data), np.array(traj2.goal_idx)
        p2.plan(traj2)
        block["parity"] = _oracle_parity_plan(_host_scene(sc), pk, n, xi0, g0, goals[:Sn], reach[:Sn],
                                              p2.history_trajectories, p2.selected_goals)
        block["parity_frac_within_1e-4"] = block["parity"]["parity_frac_within_1e-4"]
    return block


def run_config3(rank, world, scenes=100, B=256, n=30, goals_per_traj=20, streams=3, parity=True):
    streams = int(os.environ.get("OMGB_SWEEP_THREADS", streams))
    import torch

    from omg_planner_b200 import _lib
    from omg_planner_b200 import scene as S
    from omg_planner_b200.config import ChompConfig
    from omg_planner_b200.robot import PandaConstants

    dev = torch.device("cuda", torch.cuda.current_device())
    robot = PandaConstants()
    mine = [s for s in range(scenes) if s % world == rank]
    # omg/core.py:876: the -exp sweep runs without standoff
    kw = dict(goal_set_proj=True, use_standoff=False, ol_alg="MD", pre_terminate=False)
    t0 = time.perf_counter()
    plans = []
    for s in mine:
        sc = S.make_scene(num_objects=5 + s % 6, grid=128, seed=300 + s, grid_choices=[64, 96, 128], device=dev)
        cfg = ChompConfig(timesteps=n, **kw)
        goals, reach = S.make_goal_sets(B, goals_per_traj, robot.joint_lower_limit, robot.joint_upper_limit, seed=300 + s,
                                        spread=0.3)
        planner, env, traj = _goalset_planner(sc, cfg, robot, goals, reach, n)
        plans.append((planner, env, cfg, goals, sc if s == mine[0] else None, reach))
    t_build = time.perf_counter() - t0
    iters = plans[0][2].optim_steps + plans[0][2].extra_smooth_steps

    def sweep(num_threads):
        """One pass over this rank's scenes.  num_threads > 1: that many host threads, each with its own CUDA stream,
        take scenes from a shared queue (scenes are independent; every Planner owns its scene handle) -- one scene's
        host tail (history D2H, info lists) overlaps the next scenes' kernels."""
        trajs = [_fresh_traj(p, env, cfg, goals, n) for p, env, cfg, goals, _, _ in plans]
        _barrier()
        t0 = time.perf_counter()
        if num_threads <= 1:
            for k, (p, env, cfg, goals, _, _) in enumerate(plans):
                p.plan(trajs[k])
        else:
            import queue
            import threading

            todo, errs = queue.SimpleQueue(), []
            for k in range(len(plans)):
                todo.put(k)

            def worker():
                torch.cuda.set_device(dev)
                with torch.cuda.stream(torch.cuda.Stream()):
                    try:
                        while True:
                            try:
                                k = todo.get_nowait()
                            except queue.Empty:
                                return
                            plans[k][0].plan(trajs[k])
                    except Exception as e:   # noqa: BLE001
                        errs.append(e)

            ths = [threading.Thread(target=worker) for _ in range(num_threads)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            if errs:
                raise errs[0]
        torch.cuda.synchronize()
        return time.perf_counter() - t0, trajs

    sweep(1)   # warm-up
    l0 = int(_lib.lib().omgb_launch_count())
    t_seq, trajs_seq = sweep(1)
    launches = int(_lib.lib().omgb_launch_count()) - l0
    seq_final = [np.array(t.data) for t in trajs_seq]
    sweep(streams)   # warm-up of the threads' streams
    t_par, trajs_par = sweep(streams)
    same = all(np.array_equal(a, np.array(t.data)) for a, t in zip(seq_final, trajs_par))
    del seq_final, trajs_seq, trajs_par
    seq_ms = _max_over_ranks(t_seq * 1e3)
    wall_ms = _max_over_ranks(t_par * 1e3)
    total = scenes * B
    block = {
        "workload": "config3: -exp sweep, %d scenes x %d traj x %d wpt, 5-10 SDFs per scene (64^3..128^3), goal sets of %d "
                    "with online re-weighting (MD learner), no standoff (omg/core.py:876), 50 + 20 iterations, scenes "
                    "split over %d GPU(s) (%d per GPU)" % (scenes, B, n, goals_per_traj, world, len(mine)),
        "value": total * iters / (wall_ms * 1e-3), "unit": METRIC_UNIT, "scaling": "strong",
        "ms_per_step": wall_ms / (len(mine) * iters), "steps": len(mine) * iters, "scenes_per_gpu": len(mine),
        "host_threads": streams,
        "value_one_scene_at_a_time": total * iters / (seq_ms * 1e-3),
        "ms_per_step_one_scene_at_a_time": seq_ms / (len(mine) * iters),
        "concurrent_equals_sequential": bool(same),
        "timing": "host wall clock around the sweep of Planner.plan calls over this rank's scenes (numpy in / out, H2D and "
                  "D2H inside), after a warm-up sweep, max over ranks; value: %d host threads with one CUDA stream each "
                  "pull scenes from a queue (independent scenes; identical results); value_one_scene_at_a_time: a plain "
                  "loop" % streams,
        "gpu_launches_per_sweep": launches, "scene_and_planner_build_s": t_build,
    }
    if parity and rank == 0:
        Sn = 4
        planner, env, cfg, goals, sc, reach = plans[0]
        pk = dict(kw, optim_steps=10, extra_smooth_steps=4)
        pcfg = ChompConfig(timesteps=n, **pk)
        p2, env2, traj2 = _goalset_planner(sc, pcfg, robot, goals[:Sn], reach[:Sn], n)
        xi0, g0 = np.array(traj2.data), np.array(traj2.goal_idx)
        p2.plan(traj2)
        block["parity"] = _oracle_parity_plan(_host_scene(sc), pk, n, xi0, g0, goals[:Sn], reach[:Sn],
                                              p2.history_trajectories, p2.selected_goals)
        block["parity_frac_within_1e-4"] = block["parity"]["parity_frac_within_1e-4"]
    del plans
    return block


def run_all(rank, world, peak, peak_src, which=("config4", "config5", "config3"), parity=True):
    out = {}
    for name in which:
        t0 = time.perf_counter()
        try:
            if name == "config4":
                out[name] = run_config4(rank, world, peak, peak_src, parity=parity)
            elif name == "config5":
                out[name] = run_config5(rank, world, parity=parity)
            elif name == "config3":
                out[name] = run_config3(rank, world, parity=parity)
            out[name]["block_wall_s"] = time.perf_counter() - t0
        except Exception as e:   # noqa: BLE001 -- a failed block must not take the headline down with it
            import traceback
            out[name] = {"error": repr(e), "traceback": traceback.format_exc()[-1500:]}
        _barrier()
    return out


if __name__ == "__main__":
    import torch

    torch.cuda.set_device(0)
    which = tuple(sys.argv[1:]) or ("config4", "config5", "config3")
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    print(json.dumps(run_all(0, 1, peak, "measured" if os.path.exists(peaks) else "fallback", which)))
