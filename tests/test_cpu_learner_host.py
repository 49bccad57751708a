"""CPU: the host half of the Learner mirror (update rules, batched Bregman projection) against the fixtures recorded
from the reference's Learner and against the oracle."""
import glob
import os

import numpy as np
import pytest

import helpers as H
from omg_planner_b200 import scene as S
from omg_planner_b200.config import ChompConfig
from omg_planner_b200.online_learner import Learner, bregman_projection_rows
from omg_planner_b200.robot import PandaConstants
from oracle import learner_ref as LR

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "learner_*.npz")))


class _Traj(H.FakeTrajectory):
    pass


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_update_rules_on_synthetic_stream(path):
    g = np.load(path)
    alg = str(g["alg"])
    cfg = ChompConfig(ol_alg=alg)
    G = g["synthetic_cv"].shape[1]
    env = H.make_env(S.make_scene(num_objects=2, grid=16, seed=0), cfg, PandaConstants(), device="cpu")
    traj = _Traj(np.zeros((30, 9)), g["start"], g["goals"][0, 0], goal_set=list(g["goals"][0]), goal_idx=0)
    learner = Learner.__new__(Learner)
    learner.cfg, learner.env, learner.traj, learner.cost = cfg, env, traj, None
    learner._init_state(traj)
    for k, cv in enumerate(g["synthetic_cv"]):
        getattr(learner, alg)(cv)
        np.testing.assert_allclose(learner.p, g["synthetic_p"][k], rtol=1e-9, atol=1e-12)


def test_bregman_projection_rows_match_oracle():
    rng = np.random.RandomState(3)
    R_, G = 12, 9
    x = rng.dirichlet(np.ones(G), R_); v = rng.uniform(0, 0.6, (R_, G))
    delta, w = np.ones(G) / (4 * G + 1), np.ones(G)
    out = bregman_projection_rows(x, v, delta, w)
    for r in range(R_):
        np.testing.assert_allclose(out[r], LR.bregman_projection(x[r], v[r], delta, w), rtol=1e-12, atol=1e-15)




def test_two_step_bisection_equals_sequential_bit_for_bit(tmp_path):
    """omg_planner_b200/csrc/learner_bisect.h compiled for the host (tests/host/bisect_check.cpp): the two-steps-per-round
    bisection the learner kernel uses for small batches returns the sequential loop's value (find_zero,
    omg/online_learner.py:18-30) on 50 000 projections, incl. exhausted brackets, exits at the second step of a round
    and NaN stretches."""
    import os
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "bisect_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++14", "-o", exe,
                           os.path.join(here, "host", "bisect_check.cpp"), "-lm"])
    out = subprocess.run([exe, "50000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    f = dict(zip(out.stdout.split()[0::2], out.stdout.split()[1::2]))
    assert int(f["trials"]) == 50000 and int(f["mismatches"]) == 0
    assert int(f["early_exit_second"]) > 1000 and int(f["exhausted"]) > 1000 and int(f["nan_cases"]) > 100
