#!/usr/bin/env python
"""bench.py -- CHOMP trajectory-iterations/s on B200 (BASELINE.json metric), one JSON line.

  python bench.py --gpus N --steps K --warmup W          the sm_100a engine (this repo)
  python bench.py --impl reference --gpus N ...          the reference's CPU path (oracle port) on host cores

A "step" is ONE CHOMP iteration (Optimizer.optimize equivalent: cost + gradient + covariant update +
joint-limit projection) over the whole trajectory batch.  Workload at N=1: BASELINE config 2 stand-in
(1024 trajectories x 30 waypoints x 7-DOF Panda, 10 synthetic SDFs at 128^3, reference default mode:
goal-set projection with standoff, top_k_collision=1000).  N>1: every rank runs its own 1024 trajectories
against a replicated scene (weak scaling), one NCCL all-gather of final costs after the timed region.

value     = trajectories x steps / device time (CUDA events per step, inputs resident in HBM, L2 flushed
            between steps), max over ranks.
e2e       = the same through the host-buffer C-ABI entry point (omgb_chomp_step_host): pinned host xi ->
            H2D -> fused kernel -> D2H xi + info, every step.
roofline  = SURVEY 8(d) algorithmic bytes (128*P_in + 8*n*9 + 4*(2+c)*9 per trajectory-iteration, P_in
            counted by the kernel) / kernel time, against MEASURED_PEAKS.json hbm_gbs.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_MODE = dict(goal_set_proj=True, use_standoff=True, top_k_collision=1000)
FULLSUM_MODE = dict(goal_set_proj=True, use_standoff=True, top_k_collision=0)


def workload_name(a):
    return ("config2-standin: %d traj x %d wpt x 7-DOF Panda, %d synthetic SDFs @%d^3, %s"
            % (a.batch, a.waypoints, a.objects, a.grid,
               "goal-set+standoff top_k=1000 (reference defaults)" if a.mode == "default" else "goal-set+standoff full-sum"))


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's per-trajectory Python path (oracle port), one process per host core
# --------------------------------------------------------------------------------------------------
def _cpu_worker(chunk, scene, mode, n, xi, st, en, tails, barrier, total_steps):
    from oracle import chomp_ref as R

    robot = R.PandaRef()
    opts = []
    for b in chunk:
        cfg = R.RefConfig(timesteps=n, **mode)
        opts.append(R.ChompRef(robot, scene, cfg, xi[b], st[b], en[b], tails[b]))
    for _ in range(total_steps):
        barrier.wait()
        for o in opts:
            o.step()
        barrier.wait()


def run_cpu(a, scene, mode, sample, steps, warmup):
    """Returns (traj-iter/s, cores, description)."""
    from omg_planner_b200 import scene as S
    from omg_planner_b200.robot import PandaConstants

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sample = max(sample, cores)   # every host core gets at least one trajectory
    cores = max(1, min(cores, sample))
    robot = PandaConstants()
    xi, st, en, tails = S.make_trajectories(sample, a.waypoints, robot.joint_lower_limit, robot.joint_upper_limit, seed=0)
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(cores + 1)
    chunks = [list(range(w, sample, cores)) for w in range(cores)]
    procs = [ctx.Process(target=_cpu_worker, args=(chunks[w], scene, mode, a.waypoints, xi, st, en, tails, barrier,
                                                   steps + warmup), daemon=True) for w in range(cores)]
    for p in procs:
        p.start()
    total = 0.0
    for s in range(steps + warmup):
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        if s >= warmup:
            total += time.perf_counter() - t0
    for p in procs:
        p.join(timeout=30)
    value = sample * steps / total
    desc = ("%d of %d trajectories of the workload, %d processes x 1 thread, numpy oracle port of omg/cost.py + "
            "omg/optimizer.py with the C restatement of the SDF op; %d warm-up + %d timed iterations"
            % (sample, a.batch, cores, warmup, steps))
    return value, cores, desc, total / steps * 1e3


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    def __init__(self, cuda_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.h = h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2, "display_clocks": 0x100}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    mode = DEFAULT_MODE if a.mode == "default" else FULLSUM_MODE

    from omg_planner_b200 import scene as S
    from omg_planner_b200.config import ChompConfig
    from omg_planner_b200.robot import PandaConstants

    scene = S.make_scene(num_objects=a.objects, grid=a.grid, seed=0)

    # CPU baseline first (forks; must precede CUDA initialisation), rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, cores, desc, ms = run_cpu(a, scene, mode, a.cpu_sample, a.cpu_steps, 2)
        cpu = {"value": v, "unit": "trajectory-iterations/s", "cores": cores, "kind": "port", "sample": desc,
               "ms_per_step": ms}

    import torch
    import torch.distributed as dist

    from omg_planner_b200.engine import ChompEngine

    torch.cuda.set_device(local)
    numa = None
    if world > 1 and not os.environ.get("OMGB_NO_NUMA_BIND"):
        # one process per GPU: keep this rank's pinned host buffers and staging copies on the GPU's own NUMA node
        from omg_planner_b200 import dist as D0
        numa = D0.bind_host_to_device_numa(local)
        print("rank %d: numa binding %s" % (rank, numa), file=sys.stderr)
    nccl_log = None
    json_fd = None
    if world > 1:
        # NCCL logs to stdout, which must carry only the one JSON line: from here on fd 1 is stderr (NCCL's INFO lines
        # stay visible to whoever captures the run) and the JSON line is written to the saved stdout at the end.
        # Unless the caller configured NCCL's logging, INFO/INIT is switched on and also written to a per-rank file that
        # rank 0 summarises into the JSON line ("nccl": communicator size, transport lines).
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):   # (this image presets VERSION)
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
            if "NCCL_DEBUG_FILE" not in os.environ:
                import tempfile
                os.environ["NCCL_DEBUG_FILE"] = os.path.join(tempfile.gettempdir(), "omgb_nccl_n%d_%%h_%%p.log" % world)
        nccl_log = os.environ.get("NCCL_DEBUG_FILE")
        print("rank %d: NCCL_DEBUG=%s NCCL_DEBUG_SUBSYS=%s NCCL_DEBUG_FILE=%s" % (
            rank, os.environ.get("NCCL_DEBUG"), os.environ.get("NCCL_DEBUG_SUBSYS"), nccl_log), file=sys.stderr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = ChompConfig(timesteps=a.waypoints, **mode)
    robot = PandaConstants()
    eng = ChompEngine(robot=robot).load_scene(scene, cfg)
    B, n, c = a.batch, a.waypoints, cfg.constraint_rows
    xi0, st, en, tails = S.make_trajectories(B, n, robot.joint_lower_limit, robot.joint_upper_limit, seed=rank)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    xi, d_st, d_en, d_tails = dev(xi0), dev(st), dev(en), dev(tails)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    it = 0

    def one_step(timed_events=None):
        nonlocal it
        cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
        it += 1
        flush.zero_()
        if timed_events is not None:
            timed_events[0].record()
        out = eng.step(cfg, xi, d_st, d_en, d_tails)
        if timed_events is not None:
            timed_events[1].record()
        return out

    # warm-up = the timed loop's exact allocation pattern (the previous step's outputs still alive, the P_in sums):
    # a cudaMalloc by torch's caching allocator inside a timed step would stall the device for milliseconds
    wev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(a.warmup, 3))]
    wpins = []
    for k in range(len(wev)):
        out = one_step(wev[k])
        wpins.append(out["info"][:, 12].sum())
    del wev, wpins
    barrier()
    from omg_planner_b200 import _lib
    launches0 = int(_lib.lib().omgb_launch_count())
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    pins = []
    t_wall = time.perf_counter()
    for k in range(a.steps):
        out = one_step(evs[k])
        pins.append(out["info"][:, 12].sum())
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = int(_lib.lib().omgb_launch_count()) - launches0
    clocks = sampler.stop()
    dev_ms = sum(s.elapsed_time(e) for s, e in evs)
    p_in_per_launch = float(torch.stack(pins).mean().item())
    print("rank %d: device ms/step %.4f, P_in per launch %.0f" % (rank, dev_ms / a.steps, p_in_per_launch), file=sys.stderr)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())

    # warm-L2 variant (what a real plan sees: 70 back-to-back iterations, SDFs resident in L2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.steps):
        cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
        it += 1
        eng.step(cfg, xi, d_st, d_en, d_tails)
    e1.record()
    barrier()
    warm_ms = e0.elapsed_time(e1)

    # the whole fixed-goal plan as ONE persistent launch (omgb_chomp_plan: device-side queue of (iteration, trajectory)
    # items, no per-iteration launch tail).  Reported beside the per-step number, never instead of it: the SDFs stay
    # L2-resident across the iterations of a plan and there is no flush inside the launch.
    plan_iters = a.plan_iters
    xp = dev(xi0)
    eng.plan(cfg, xp, d_st, d_en, d_tails, iters=plan_iters)   # warm-up
    xp.copy_(dev(xi0))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.plan(cfg, xp, d_st, d_en, d_tails, iters=plan_iters)
    e1.record()
    barrier()
    plan_ms = e0.elapsed_time(e1)

    # end to end through the host-buffer C-ABI call (omgb_chomp_step_host), pinned host memory.  Three transfer
    # strategies of the same call are timed; the headline e2e is the library default (mode 0).
    h_xi = torch.from_numpy(xi0.copy()).pin_memory()
    h_st, h_en, h_tails = (torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (st, en, tails))
    n_xi, n_se, n_goal, n_info = B * n * 9, B * 9, B * c * 9, B * 16
    it_h = 0
    hx, hs, he, ht = h_xi.numpy(), h_st.numpy(), h_en.numpy(), h_tails.numpy()   # the caller's numpy views

    def host_step(events=None):
        nonlocal it_h
        cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it_h + 1)
        it_h += 1
        flush.zero_()
        if events is not None:
            # the flush must be over before the timed region starts: otherwise the host side of the call (argument
            # marshalling, launch) would run hidden under it and the events would only see the kernel + sync
            torch.cuda.synchronize()
            events[0].record()
        eng.step_host(cfg, hx, hs, he, ht)
        if events is not None:
            events[1].record()

    e2e_modes = {}
    for mode_id, mode_name in ((1, "staged"), (2, "staged_pipelined"), (0, "default_zero_copy")):
        eng.set_host_mode(mode_id)
        h_xi.copy_(torch.from_numpy(xi0))
        it_h = 0
        for _ in range(3):
            host_step()
        barrier()
        hevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        for k in range(a.steps):
            host_step(hevs[k])
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in hevs)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_modes[mode_name] = float(t.item())
    e2e_ms_max = e2e_modes["default_zero_copy"]

    # the one collective of the path: all-gather of the final per-trajectory costs, timed on the device
    final_cost = out["info"][:, 2].contiguous()
    allgather_ms = None
    if world > 1:
        from omg_planner_b200 import dist as D
        D.all_gather_costs(final_cost)   # warm-up: communicator channels, buffers
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gathered = D.all_gather_costs(final_cost)
        e1.record()
        torch.cuda.synchronize()
        assert gathered.shape[0] == world * B
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allgather_ms = float(t.item())

    # the reference's plugin call: Optimizer.optimize(traj, force_update=True) with a host numpy trajectory, one
    # trajectory (the reference's own shape) and the whole batch; wall clock per call incl. H2D / D2H and Python
    e2e_plugin = None
    if not a.no_plugin:
        e2e_plugin = run_plugin_e2e(a, scene, mode, xi0, st, en, tails, robot, steps=a.steps)
        t = torch.tensor([e2e_plugin["batch"]["ms_per_call"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_plugin["batch"]["ms_per_call_max_over_ranks"] = float(t.item())
        e2e_plugin["batch"]["value"] = B * world / (float(t.item()) * 1e-3)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # BASELINE configs 3 / 4 / 5 at full size (every rank takes part: they shard over WORLD_SIZE)
    configs = None
    if a.configs:
        del flush
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        which = tuple("config" + c.strip() for c in a.configs.split(",") if c.strip())
        configs = bench_configs.run_all(rank, world, peak, peak_src, which, parity=not a.no_parity)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    bytes_per_launch = 128.0 * p_in_per_launch + B * (8.0 * n * 9 + 4.0 * (2 + c) * 9)
    kern_ms = dev_ms / a.steps
    achieved = bytes_per_launch / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("%s:flush:%dx%dx%dx%d" % (a.mode, a.batch, a.waypoints, a.objects, a.grid))
    total = B * world * a.steps
    line = {
        "metric": "CHOMP trajectory-iterations/s (batch traj x waypt)", "value": total / (dev_ms_max * 1e-3),
        "unit": "trajectory-iterations/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 state/FK/gradient, f32 SDF operator (as the reference)", "data": "synthetic",
        "config": {"workload": workload_name(a), "batch_per_gpu": B, "waypoints": n, "objects": a.objects,
                   "grid": a.grid, "parallelism": "dp%d (trajectory batch sharded, scene replicated)" % world,
                   "l2": "flushed between timed steps (256 MiB write); value_warm_l2 = back-to-back steps",
                   "timing": "CUDA events around each step on the launch stream, summed; max over ranks"},
        "value_warm_l2_rank0": B * a.steps / (warm_ms * 1e-3),
        "plan_persistent_rank0": {"value": B * plan_iters / (plan_ms * 1e-3), "unit": "trajectory-iterations/s",
                                  "iterations": plan_iters, "ms_per_iteration": plan_ms / plan_iters,
                                  "api": "omgb_chomp_plan: one persistent launch for the whole fixed-goal plan "
                                         "(reference schedule, 50 + 20 iterations), L2 warm, no flush inside"},
        "e2e": {"value": total / (e2e_ms_max * 1e-3), "unit": "trajectory-iterations/s",
                "h2d_bytes_per_step": 8 * (n_xi + 2 * n_se + n_goal), "d2h_bytes_per_step": 8 * (n_xi + n_info),
                "ms_per_step": e2e_ms_max / a.steps,
                "api": "omgb_chomp_step_host (pinned host buffers, mode 0: the fused kernel reads xi/start/end/goal "
                       "rows from and writes xi/info to mapped pinned host memory over PCIe -- the H2D/D2H bytes move "
                       "inside the kernel; stream synchronised before return); timed per step by CUDA events recorded "
                       "on an idle device (synchronised after the L2 flush), so the host side of the call is inside",
                "ms_per_step_by_transfer_mode": {k: v / a.steps for k, v in e2e_modes.items()}},
        "e2e_plugin": e2e_plugin,
        "allgather_ms": allgather_ms,
        "allgather": None if allgather_ms is None else "NCCL all_gather_into_tensor of %d fp64 final costs per rank, "
                                                       "CUDA events, max over ranks; outside the timed steps" % B,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "chomp_step_kernel<%s>" % ("topk" if a.mode == "default" else "fullsum"),
                     "algorithmic_bytes_per_launch": bytes_per_launch, "p_in_per_launch": p_in_per_launch,
                     "kernel_ms": kern_ms,
                     "note": "algorithmic bytes per SURVEY 8d (128 B x in-bounds pairs + state); exact culling keeps the "
                             "measured DRAM traffic far below it -- see DESIGN.md section 6"},
        "clocks": clocks, "wall_s_timed_region": t_wall,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if configs is not None:
        line["configs"] = configs
    if world > 1:
        line["nccl"] = nccl_summary(nccl_log, world)
        line["numa_binding_rank0"] = numa
    if world == 1 and not a.no_aux:
        # the kernels either side of the CHOMP loop (goal-set IK, SDF packing, point-cloud field, trajectory
        # initialisation): reported beside the headline, never part of it
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_aux
            line["aux_kernels"] = bench_aux.run_aux(peak)
        except Exception as e:   # noqa: BLE001
            line["aux_kernels"] = {"error": repr(e)}
    if json_fd is not None:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def nccl_summary(pattern, world):
    """Rank 0's view of the NCCL INFO log files: communicator size and the transport lines, echoed to stderr."""
    import glob
    import re
    import socket

    out = {"debug_file": pattern, "nranks": None, "lines": [], "torch_world_size": world}
    try:
        import torch
        out["nccl_version"] = ".".join(str(v) for v in torch.cuda.nccl.version())
    except Exception:   # noqa: BLE001
        pass
    if not pattern:
        return out
    path = pattern.replace("%h", socket.gethostname()).replace("%p", "*")
    for f in sorted(glob.glob(path)):
        try:
            for ln in open(f, errors="replace"):
                m = re.search(r"nranks (\d+)", ln)
                if m and out["nranks"] is None:
                    out["nranks"] = int(m.group(1))
                if ("nranks" in ln or "NVLS" in ln or "NCCL version" in ln) and len(out["lines"]) < 12:
                    out["lines"].append(ln.strip()[:200])
        except OSError:
            pass
    for ln in out["lines"]:
        print("NCCL:", ln, file=sys.stderr)
    out["comm_nranks_ok"] = (out["nranks"] == world) if out["nranks"] is not None else None
    return out


def run_plugin_e2e(a, scene, mode, xi0, st, en, tails, robot, steps):
    """Optimizer.optimize(traj, force_update=True) -- the call omg/planner.py:621 makes -- through the plugin classes:
    host numpy trajectory in, updated trajectory + info out, every call (pinned staging, H2D, fused kernel, D2H)."""
    import types

    import torch

    from omg_planner_b200.config import ChompConfig
    from omg_planner_b200.cost import Cost
    from omg_planner_b200.optimizer import Optimizer

    cfg = ChompConfig(timesteps=a.waypoints, **mode)
    env = types.SimpleNamespace()
    env.config = cfg
    env.target_idx = scene["target_idx"]
    env.objects = [types.SimpleNamespace(name=nm, pose_mat=np.array(scene["pose_mats"][i]), attached=False,
                                         reach_grasps=[], grasps=[]) for i, nm in enumerate(scene["names"])]
    env.sdf_torch = torch.from_numpy(scene["sdf_grids"]).cuda()
    env.sdf_limits = torch.from_numpy(scene["sdf_limits"]).cuda()
    rk = types.SimpleNamespace(_pose_0=robot.pose_0, _tip2joint=robot.tip2joint, _joint_axis=robot.joint_axis,
                               _joint_origin=robot.joint_axis, center_offset=robot.center_offset)
    env.robot = types.SimpleNamespace(robot_kinematics=rk, collision_points=robot.collision_points,
                                      joint_lower_limit=robot.joint_lower_limit, joint_upper_limit=robot.joint_upper_limit)

    class Traj(object):   # omg/core.py:23-57 carrier
        def __init__(self, data, start, end, goal_idx):
            self.data, self.start, self.end, self.goal_set, self.goal_idx = data, start, end, [], goal_idx

        def set(self, x):
            self.data = x

    cost = Cost(env)
    target = env.objects[env.target_idx]
    cost.target_obj = target
    res = {}
    B = xi0.shape[0]
    for name, traj in (("single", Traj(xi0[0].copy(), st[0], en[0], 0)),
                       ("batch", Traj(xi0.copy(), st, en, np.zeros(B, dtype=int)))):
        if mode["goal_set_proj"]:
            target.reach_grasps = [tails[0]] if name == "single" else tails[:, None]
        opt = Optimizer(env, cost)
        calls = max(steps, 20) if name == "single" else steps
        for _ in range(6):   # (past the transient: the result arrays cycle through three pinned blocks)
            opt.optimize(traj, force_update=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(calls):
            info = opt.optimize(traj, force_update=True)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / calls * 1e3
        nb = 1 if name == "single" else B
        res[name] = {"trajectories": nb, "ms_per_call": ms, "value": nb / (ms * 1e-3), "unit": "trajectory-iterations/s",
                     "calls": calls}
    res["api"] = ("omg_planner_b200.optimizer.Optimizer.optimize(traj, force_update=True) over Cost.evaluate: numpy "
                  "trajectory -> H2D (straight from the array when it is the pinned result of the previous call, else "
                  "through pinned staging) -> ONE fused launch -> D2H of xi into a fresh pinned array (the new "
                  "traj.data) + the [B,16] info array; gradient / cost_traj / per-trajectory dicts fetched lazily for "
                  "a batch; host wall clock per call")
    res["reference_call"] = "omg/planner.py:621 self.optim.optimize(traj, force_update=True) (omg/optimizer.py:115-135)"
    return res


def run_reference(a):
    """The reference arm: the reference's own CPU implementation of the path (numpy Python + C operator;
    the oracle port, because /root/reference does not travel to the GPU box) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from omg_planner_b200 import scene as S

    mode = DEFAULT_MODE if a.mode == "default" else FULLSUM_MODE
    scene = S.make_scene(num_objects=a.objects, grid=a.grid, seed=0)
    sample = a.cpu_sample
    steps, warmup = max(1, a.steps), max(1, min(a.warmup, 3))
    v, cores, desc, ms = run_cpu(a, scene, mode, sample, steps, warmup)
    line = {
        "impl": "reference", "metric": "CHOMP trajectory-iterations/s (batch traj x waypt)", "value": v,
        "unit": "trajectory-iterations/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 numpy + f32 SDF operator",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "batch_per_gpu": a.batch, "waypoints": a.waypoints,
                   "objects": a.objects, "grid": a.grid, "sample_trajectories_per_step": sample},
        "cpu_baseline": {"value": v, "unit": "trajectory-iterations/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "trajectory-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--waypoints", type=int, default=30)
    ap.add_argument("--objects", type=int, default=10)
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--mode", default="default", choices=["default", "fullsum"])
    ap.add_argument("--plan-iters", type=int, default=70)
    ap.add_argument("--cpu-sample", type=int, default=64)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--no-plugin", action="store_true", help="skip the Optimizer.optimize plugin-call timing")
    ap.add_argument("--configs", default="4,5,3", help="BASELINE configs to run beside the config-2 headline "
                                                       "(comma list of 3,4,5; empty string: none)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle-checked subsamples of the config blocks")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
