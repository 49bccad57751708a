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


