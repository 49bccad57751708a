"""CPU: oracle/learner_ref.py (goal scoring + goal-distribution update) against the fixtures recorded from the
reference's own Learner (tools/make_golden_learner.py)."""
import glob
import os

import numpy as np
import pytest

from omg_planner_b200 import scene as S
from oracle import chomp_ref as R
from oracle import learner_ref as LR

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "learner_*.npz")))


def test_fixtures_present():
    assert len(GOLDEN) >= 5


def replay(g, cost_fn, iters=None, trajs=None):
    """The Planner.plan interleave (omg/planner.py:612-621) with `cost_fn(b, xi, t) -> cost vector` supplying the
    Learner's cost vectors and the oracle doing the CHOMP step.  Yields per-iteration records."""
    alg, standoff = str(g["alg"]), bool(int(g["use_standoff"]))
    sc = S.make_scene(**eval(str(g["scene_args"])))
    robot = R.PandaRef(body_points=g["body_points"])
    goals, reach, start = g["goals"], g["reach"], g["start"]
    iters = g["history"].shape[1] - 1 if iters is None else iters
    for b in (range(goals.shape[0]) if trajs is None else trajs):
        cfg = R.RefConfig(goal_set_proj=True, use_standoff=standoff, top_k_collision=1000, ol_alg=alg)
        learner = LR.LearnerRef(cfg, goals.shape[1])
        cv0 = cost_fn(b, sc, robot, cfg, None, 0.0)
        idx = int(np.argmin(cv0))                                    # omg/online_learner.py:95-102
        xi = S.clamped_cubic(start, goals[b, idx], cfg.timesteps)
        yield b, 0, cost_fn(b, sc, robot, cfg, xi, 0.0), learner.p.copy(), idx, xi   # (fixture: scored after re-init)
        opt = R.ChompRef(robot, sc, cfg, xi, start, goals[b, idx],
                         reach[b, idx] if standoff else goals[b, idx][None])
        for it in range(iters):
            learner.t += 1
            cv = cost_fn(b, sc, robot, cfg, opt.xi, learner.t)
            idx = learner.update(cv)
            opt.end = goals[b, idx].copy(); opt.goal = goals[b, idx].copy()
            opt.goal_rows = np.atleast_2d(reach[b, idx] if standoff else goals[b, idx][None])
            opt.step()
            yield b, it + 1, cv, learner.p.copy(), idx, opt.xi.copy()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_oracle_learner_matches_reference(path):
    g = np.load(path)
    standoff = bool(int(g["use_standoff"]))
    goals, reach, start = g["goals"], g["reach"], g["start"]

    def cost_fn(b, sc, robot, cfg, xi, t):
        if xi is None:   # Learner.__init__ scores from the initial interpolation to goal 0
            xi = S.clamped_cubic(start, goals[b, 0], cfg.timesteps)
        rg = reach[b][:, -1, :] if standoff else goals[b]
        return LR.cost_vector(robot, sc, cfg, xi, goals[b], rg, t)

    for b, it, cv, p, idx, xi in replay(g, cost_fn, iters=6, trajs=[0, 2]):
        np.testing.assert_allclose(cv, g["cost_vectors"][b, it], rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(p, g["p"][b, it], rtol=1e-6, atol=1e-9)
        assert idx == g["selected"][b, it]
        assert np.abs(xi - g["history"][b, it])[:, :7].max() <= 1e-9


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_update_rule_on_synthetic_cost_stream(path):
    g = np.load(path)
    cfg = R.RefConfig(ol_alg=str(g["alg"]))
    learner = LR.LearnerRef(cfg, g["synthetic_cv"].shape[1])
    for k, cv in enumerate(g["synthetic_cv"]):
        learner.update(cv)
        np.testing.assert_allclose(learner.p, g["synthetic_p"][k], rtol=1e-9, atol=1e-12)


def test_interpolation_and_first_waypoint():
    from scipy import interpolate

    rng = np.random.RandomState(0)
    start, goals = rng.randn(9), rng.randn(4, 9)
    for n in (1, 7, 30):
        f = interpolate.interp1d(np.linspace(0, 1, 2), np.stack([np.tile(start, (4, 1)), goals]), "linear", axis=0)
        ref = np.transpose(f(np.linspace(0, 1, n + 2)[1:-1]), (1, 0, 2)).reshape(-1, 9)
        np.testing.assert_array_equal(LR.interpolate_to_goals(start, goals, n), ref)
    assert [LR.first_waypoint(t, 50, 30) for t in (0.0, 1.0, 2.0, 25.0, 49.0, 50.0, 70.0)] == [0, 0, 1, 15, 29, 29, 29]
