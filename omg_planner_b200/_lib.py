"""ctypes binding of libomgb200.so (include/omgb200.h).  There is NO fallback: if the CUDA library is
missing or fails to load, importing the product fails loudly."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("OMGB_LIB") or os.path.join(_HERE, "lib", "libomgb200.so")   # OMGB_LIB: A/B experiments
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("omgb200.cu", "chomp_kernels.cuh", "goal_kernels.cuh",
                                                     "sdf_device.cuh", "sdf_asset_kernels.cuh", "traj_kernels.cuh", "learner_kernels.cuh", "learner_bisect.h",
                                                     "ik_kernels.cu", "ik_svd_reg.cuh", "host_common.h")] + [
    os.path.join(ROOT, "include", "omgb200.h")]

INFO_STRIDE = 16
INFO_KEYS = ["obs", "smooth", "cost", "collide", "reach", "grad", "weighted_obs_grad", "weighted_smooth_grad",
             "terminate", "violate_limit", "execute", "failure_terminate", "p_in", "nonzero", "limit_rounds",
             "reserved"]

EXPORTS = ["omgb_version", "omgb_last_error", "omgb_scene_create", "omgb_scene_destroy", "omgb_scene_set_robot",
           "omgb_scene_set_sdf", "omgb_scene_set_sdf_layout", "omgb_scene_set_profile", "omgb_scene_set_options", "omgb_scene_set_host_mode", "omgb_launch_count", "omgb_scene_set_objects", "omgb_scene_set_metric",
           "omgb_sdf_loss_workspace_bytes", "omgb_sdf_loss", "omgb_chomp_step", "omgb_chomp_plan",
           "omgb_chomp_step_host", "omgb_batch_obstacle_cost", "omgb_goal_costs", "omgb_chomp_plan_history",
           "omgb_traj_interpolate", "omgb_sdf_pack", "omgb_point_sdf", "omgb_ik_solve", "omgb_hand_poses", "omgb_chomp_plan_step",
           "omgb_learner_update", "omgb_chomp_plan_goalset"]


class StepParams(ctypes.Structure):
    """omgb_step_params_t"""
    _fields_ = [
        ("n_waypoints", ctypes.c_int32), ("goal_set_proj", ctypes.c_int32), ("constraint_rows", ctypes.c_int32),
        ("top_k_collision", ctypes.c_int32), ("uncheck_finger_collision", ctypes.c_int32),
        ("consider_finger", ctypes.c_int32), ("allow_collision_point", ctypes.c_int32),
        ("pre_terminate", ctypes.c_int32), ("joint_limit_max_steps", ctypes.c_int32), ("update", ctypes.c_int32),
        ("time_interval", ctypes.c_double), ("obstacle_weight", ctypes.c_double),
        ("smoothness_weight", ctypes.c_double), ("step_size", ctypes.c_double), ("clip_grad_scale", ctypes.c_double),
        ("terminate_smooth_loss", ctypes.c_double), ("link_smooth_weight", ctypes.c_double * 9),
    ]


LEARNER_ALGS = {"FTL": 0, "FTC": 1, "Exp": 2, "MD": 3, "Proj": 4, "INIT": 5}


class LearnerParams(ctypes.Structure):
    """omgb_learner_params_t"""
    _fields_ = [("alg", ctypes.c_int32), ("num_goals", ctypes.c_int32), ("n_waypoints", ctypes.c_int32),
                ("first_waypoint", ctypes.c_int32), ("constraint_rows", ctypes.c_int32),
                ("normalize_cost", ctypes.c_int32), ("base_obstacle_weight", ctypes.c_double),
                ("smoothness_base_weight", ctypes.c_double), ("dist_eps", ctypes.c_double), ("eta", ctypes.c_double),
                ("etas", ctypes.c_double * 5)]


class GoalsetPlanBuffers(ctypes.Structure):
    """omgb_goalset_plan_buffers_t"""
    _fields_ = [(k, ctypes.c_void_p) for k in (
        "xi", "start", "end", "goal_rows", "done", "info", "hist_xi", "hist_info", "goal_set", "reach", "reach_goals",
        "p", "sum_costs", "experts_p", "experts_costs", "q", "goal_idx", "selected", "collision")] + [
        ("goals_shared", ctypes.c_int32), ("reserved_", ctypes.c_int32)]


class SdfSource(ctypes.Structure):
    """omgb_sdf_source_t"""
    _fields_ = [("data", ctypes.c_void_p), ("shape", ctypes.c_int32 * 3), ("layout", ctypes.c_int32),
                ("dtype", ctypes.c_int32), ("scale", ctypes.c_float)]


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler",
              "-fPIC,-ffp-contract=off"]
# translation units: (source, extra flags).  ik_kernels.cu is compiled without FMA contraction so that the Newton
# iteration rounds like the CPU code it is checked against.
UNITS = [("omgb200.cu", []), ("ik_kernels.cu", ["-fmad=false"])]


def nvcc_commands(out=LIB_PATH, verbose=False):
    obj_dir = os.path.join(os.path.dirname(out), "obj")
    cmds, objs = [], []
    for src, extra in UNITS:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmds.append(["nvcc"] + (["-Xptxas=-v"] if verbose else []) + NVCC_FLAGS + extra +
                    ["-c", "-o", obj, os.path.join(_HERE, "csrc", src)])
        objs.append(obj)
    cmds.append(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs)
    return obj_dir, cmds


def build(force=False, verbose=False):
    """Compile libomgb200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    newest = max(os.path.getmtime(s) for s in SOURCES)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    obj_dir, cmds = nvcc_commands(verbose=verbose)
    os.makedirs(obj_dir, exist_ok=True)
    procs = [subprocess.Popen(c) for c in cmds[:-1]]
    for p, c in zip(procs, cmds[:-1]):
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, c)
    subprocess.check_call(cmds[-1])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libomgb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the CHOMP hot path." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    L.omgb_version.restype = ci
    L.omgb_last_error.restype = ctypes.c_char_p
    L.omgb_scene_create.argtypes = [ctypes.POINTER(vp), ci]
    L.omgb_scene_destroy.argtypes = [vp]
    L.omgb_scene_set_robot.argtypes = [vp, vp, vp, vp, vp, ci, vp, vp, ci, vp, vp]
    L.omgb_scene_set_sdf.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp]
    L.omgb_scene_set_sdf_layout.argtypes = [vp, ci, vp]
    L.omgb_scene_set_profile.argtypes = [vp, vp]
    L.omgb_scene_set_options.argtypes = [vp, ci, ci]
    L.omgb_scene_set_host_mode.argtypes = [vp, ci]
    L.omgb_launch_count.restype = ctypes.c_ulonglong
    L.omgb_scene_set_objects.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.omgb_scene_set_metric.argtypes = [vp, ci, vp, ci, vp]
    L.omgb_sdf_loss_workspace_bytes.restype = ctypes.c_size_t
    L.omgb_sdf_loss_workspace_bytes.argtypes = [ci]
    L.omgb_sdf_loss.argtypes = [vp] * 8 + [ci] * 5 + [vp] * 5
    L.omgb_chomp_step.argtypes = [vp, ctypes.POINTER(StepParams), ci] + [vp] * 11
    L.omgb_chomp_plan.argtypes = [vp, ctypes.POINTER(StepParams), ci, vp, vp, vp, ci, ci] + [vp] * 7
    L.omgb_chomp_plan_history.argtypes = [vp, ctypes.POINTER(StepParams), ci, vp, vp, vp, ci, ci] + [vp] * 9
    L.omgb_traj_interpolate.argtypes = [vp, ci, ci, ci, ci, vp, vp]
    L.omgb_sdf_pack.argtypes = [ctypes.POINTER(SdfSource), ci, ci, ci, ci, vp, vp]
    L.omgb_point_sdf.argtypes = [vp, ci, vp, vp, vp, ci, ci, ci, vp, vp, vp]
    L.omgb_ik_solve.argtypes = [vp, vp, vp, vp, ci, ci, vp, ci, vp, vp, vp, vp]
    L.omgb_hand_poses.argtypes = [vp, vp, ctypes.c_longlong, ci, vp, vp]
    L.omgb_chomp_plan_step.argtypes = [vp, ctypes.POINTER(StepParams), ci, ci, ci] + [vp] * 9
    L.omgb_learner_update.argtypes = [ctypes.POINTER(LearnerParams), ci, vp, vp, vp, ci] + [vp] * 13
    L.omgb_chomp_plan_goalset.argtypes = [vp, ctypes.POINTER(StepParams), ctypes.POINTER(LearnerParams), ci, ci, vp, vp, ci,
                                          ctypes.POINTER(GoalsetPlanBuffers), cd, ctypes.POINTER(ci), vp]
    L.omgb_chomp_step_host.argtypes = [vp, ctypes.POINTER(StepParams), ci] + [vp] * 6
    L.omgb_batch_obstacle_cost.argtypes = [vp, vp, ci, ci, vp, cd, ci, vp, vp, vp, vp]
    L.omgb_goal_costs.argtypes = [vp, ci, vp, ctypes.c_longlong, vp, ci, ci, ci, cd, ci, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("omgb_version", "omgb_last_error", "omgb_sdf_loss_workspace_bytes", "omgb_launch_count"):
            fn.restype = ci
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().omgb_last_error()
        raise RuntimeError("%s failed (status %d): %s" % (what or "libomgb200", rc, msg.decode() if msg else "?"))
