"""CPU: host-side logic of the product and the C-ABI surface (no GPU compute)."""
import os
import re
import subprocess

import numpy as np
import pytest

from omg_planner_b200 import _lib
from omg_planner_b200 import scene as S
from omg_planner_b200.config import ChompConfig
from omg_planner_b200.robot import PandaConstants
from oracle import chomp_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    _lib.build()
    L = _lib.lib()
    assert L.omgb_version() == 100
    header = open(os.path.join(ROOT, "include", "omgb200.h")).read()
    declared = set(re.findall(r"\b(omgb_[a-z_0-9]+)\s*\(", header))
    declared -= {"omgb_scene", "omgb_status"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name


def test_header_is_plain_c(tmp_path):
    """include/omgb200.h must compile as C99 (the boundary is a C ABI: no C++, no torch types) and its struct layouts
    must match the ctypes mirrors."""
    import ctypes
    import subprocess

    src = tmp_path / "h.c"
    src.write_text('#include <stdio.h>\n#include "omgb200.h"\nint main(void){printf("%zu %zu %zu\\n", '
                   'sizeof(omgb_step_params_t), sizeof(omgb_learner_params_t), sizeof(omgb_sdf_source_t));'
                   'printf("%zu %zu %zu\\n", sizeof(omgb_goalset_plan_buffers_t), '
                   'offsetof(omgb_goalset_plan_buffers_t, collision), offsetof(omgb_goalset_plan_buffers_t, goals_shared));'
                   'return 0;}\n')
    exe = tmp_path / "h"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-o", str(exe), str(src)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(_lib.StepParams), ctypes.sizeof(_lib.LearnerParams), ctypes.sizeof(_lib.SdfSource),
                     ctypes.sizeof(_lib.GoalsetPlanBuffers), _lib.GoalsetPlanBuffers.collision.offset,
                     _lib.GoalsetPlanBuffers.goals_shared.offset]


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.omgb_scene_set_robot(None, None, None, None, None, 0, None, None, 15, None, None) == -1
    assert b"null" in L.omgb_last_error()
    assert L.omgb_sdf_loss_workspace_bytes(10) > 0
    assert L.omgb_batch_obstacle_cost(None, None, 0, 0, None, 0.1, 0, None, None, None, None) == -1


def test_argument_validation_of_the_data_path_entry_points_without_gpu():
    """Entry points added around the CHOMP loop reject bad arguments before touching CUDA (status -1, message set)."""
    import ctypes

    L = _lib.lib()
    assert L.omgb_traj_interpolate(None, 4, 1, 30, 1, None, None) == -1          # fewer than 2 knots
    assert L.omgb_traj_interpolate(None, 4, 2, 30, 7, None, None) == -1          # unknown mode
    assert L.omgb_traj_interpolate(None, 0, 2, 30, 1, None, None) == 0           # empty batch is fine
    assert L.omgb_sdf_pack(None, 65, 8, 8, 8, None, None) == -1                  # more than OMGB_MAX_OBJECTS
    assert L.omgb_sdf_pack(None, 0, 8, 8, 8, None, None) == 0
    assert L.omgb_point_sdf(None, 0, None, None, None, 4, 4, 4, None, None, None) == -1   # no points
    frames = np.tile(np.eye(4), (8, 1, 1))
    lim = np.zeros(7)
    assert L.omgb_ik_solve(frames.ctypes.data, lim.ctypes.data, lim.ctypes.data, None, 3, 0, None, 2, None, None, None,
                           None) == -1                                             # chain_length < 1
    assert L.omgb_ik_solve(frames.ctypes.data, None, None, None, 3, 1, None, 2, None, None, None, None) == -1
    assert L.omgb_ik_solve(frames.ctypes.data, lim.ctypes.data, lim.ctypes.data, None, 0, 1, None, 2, None, None, None,
                           None) == 0                                              # no poses
    assert L.omgb_hand_poses(frames.ctypes.data, None, 3, 5, None, None) == -1   # stride < 7
    prm = _lib.LearnerParams()
    prm.alg, prm.num_goals, prm.n_waypoints, prm.first_waypoint, prm.constraint_rows = 9, 4, 30, 0, 1
    assert L.omgb_learner_update(ctypes.byref(prm), 2, None, None, None, 0, *([None] * 13)) == -1   # unknown algorithm
    prm.alg, prm.num_goals = 3, 300
    assert L.omgb_learner_update(ctypes.byref(prm), 2, None, None, None, 0, *([None] * 13)) == -1   # > 256 goals
    assert b"256" in L.omgb_last_error()
    ran = ctypes.c_int(7)                                                          # whole goal-set plan: null scene
    assert L.omgb_chomp_plan_goalset(None, None, None, 3, 1, None, None, 2, None, -1.0, ctypes.byref(ran), None) == -1
    assert ran.value == 0


@pytest.mark.parametrize("gsp", [True, False])
@pytest.mark.parametrize("n", [30, 50, 7])
def test_metric_matrices_match_oracle_and_closed_form(gsp, n):
    cfg = ChompConfig(goal_set_proj=gsp, timesteps=30)
    cfg.get_global_param(n)
    ref = R.RefConfig(goal_set_proj=gsp, timesteps=30)
    ref.set_timesteps(n)
    np.testing.assert_array_equal(cfg.diff_matrices[0], ref.K[0])
    np.testing.assert_array_equal(cfg.diff_matrices[1], ref.K[1])
    np.testing.assert_array_equal(cfg.Ainv, ref.Ainv)
    assert cfg.time_interval == ref.time_interval
    # SURVEY Appendix B closed form of the metric inverse
    i = np.arange(1, n + 1)
    mn, mx = np.minimum.outer(i, i), np.maximum.outer(i, i)
    closed = cfg.time_interval ** 2 * (mn if gsp else mn * (n + 1 - mx) / (n + 1))
    np.testing.assert_allclose(cfg.Ainv, closed, rtol=1e-9, atol=1e-12)


def test_projection_matrix_and_schedule():
    cfg = ChompConfig()
    M = cfg.projection_matrix()
    n, c = 30, 5
    C = np.zeros([c, n]); C[-c:, -c:] = np.eye(c)
    np.testing.assert_allclose(C.dot(M), np.eye(c), atol=1e-9)   # C M = I: the projected update satisfies C xi = goal
    assert cfg.constraint_rows == 5 and ChompConfig(use_standoff=False).constraint_rows == 1
    assert ChompConfig(goal_set_proj=False).constraint_rows == 0
    ow, sw, ss = cfg.schedule(3)
    assert ow == 1.0 and sw == 0.1 * 1.02 ** 3 and ss == 0.1


def test_clamped_cubic_matches_scipy():
    from scipy import interpolate
    a, b = S.START_CONF, S.START_CONF + np.linspace(-1, 1, 9)
    f = interpolate.CubicSpline(np.linspace(0, 1, 2), np.stack([a, b]), bc_type="clamped")
    np.testing.assert_allclose(S.clamped_cubic(a, b, 30), f(np.linspace(0, 1, 32)[1:-1]), atol=1e-12)


def test_scene_layout_follows_combine_sdfs():
    sc = S.make_scene(num_objects=4, grid=40, seed=1, grid_choices=[24, 32, 40])
    g, lim = sc["sdf_grids"], sc["sdf_limits"]
    assert g.shape[1:] == (40, 40, 40) and g.dtype == np.float32
    for i in range(4):
        assert tuple(lim[i, 6:9]) == (40, 40, 40)
        np.testing.assert_allclose((lim[i, 3:6] - lim[i, 0:3]) / 40, lim[i, 9], rtol=1e-5)   # g = (q-min)/delta
    small = [i for i in range(3) if (g[i, 39] == 1.0).all()]
    assert small, "at least one object should be padded with 1.0"
    rob = PandaConstants()
    xi, st, en, tails = S.make_trajectories(4, 30, rob.joint_lower_limit, rob.joint_upper_limit)
    assert (tails[:, -1] == en).all() and (en[:, :7] > rob.joint_lower_limit[0, :7]).all()


def test_step_kernel_shared_memory_layout(tmp_path):
    """The fused step's CTA shapes rest on footprints: three CTAs per SM for 30 waypoints, two for 50-60
    (omg_planner_b200/csrc/omgb200.cu: launch_step).  Host-compiled make_layout: regions ordered and sized, shapes fit."""
    exe = str(tmp_path / "step_layout_check")
    subprocess.check_call(["nvcc", "-std=c++17", "-arch=sm_100a", "-o", exe, os.path.join(ROOT, "tests", "host", "step_layout_check.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = [dict(zip(ln.split()[0::2], ln.split()[1::2])) for ln in out.stdout.strip().splitlines()]
    assert len(rows) == 7 and all(r["bad"] == "0" for r in rows), out.stdout
    for r in rows[:5]:
        assert r["fits"] == "1", r
