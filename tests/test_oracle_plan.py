"""CPU: the oracle's outer loop (oracle/planner_ref.py) replayed against fixtures recorded from the reference's own
Planner.plan (tools/make_golden_plan.py -> tests/golden/plan_*.npz)."""
import glob
import os

import numpy as np
import pytest

from omg_planner_b200 import scene as S
from oracle import chomp_ref as R
from oracle import learner_ref as LR
from oracle import planner_ref as P

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "plan_*.npz")))


def run_oracle_plan(g, sc, b, robot):
    gsp, standoff, alg = bool(g["goal_set_proj"]), bool(g["use_standoff"]), str(g["ol_alg"])
    cfg = R.RefConfig(goal_set_proj=gsp, use_standoff=standoff, top_k_collision=1000, ol_alg=alg,
                      optim_steps=int(g["optim_steps"]), extra_smooth_steps=int(g["extra_smooth_steps"]),
                      pre_terminate=bool(g["pre_terminate"]))
    learner_on = gsp and alg not in ("Baseline", "Proj")
    if not learner_on:
        rows = None
        if gsp:
            rows = g["tails"][b] if standoff else g["end"][b][None]
        return P.plan(robot, sc, cfg, g["xi0"][b], g["start"][b], g["end"][b], rows)
    goals, reach = g["goals"][b], g["reach"][b]
    # Learner.__init__ (omg/online_learner.py:91-102): initial goal = argmin of the cost vector at t = 0, then the
    # trajectory is re-initialised towards it -- the fixture's xi0 is that state
    learner = LR.LearnerRef(cfg, goals.shape[0])
    init = np.zeros((cfg.timesteps, 9))
    from oracle import traj_ref as T
    init = T.interpolate_waypoints(np.stack([g["start"][b], goals[0]]), cfg.timesteps)
    cv0 = LR.cost_vector(robot, sc, cfg, init, goals, reach[:, -1, :] if standoff else goals, 0.0)
    g0 = int(np.argmin(cv0))
    xi0 = T.interpolate_waypoints(np.stack([g["start"][b], goals[g0]]), cfg.timesteps)
    np.testing.assert_allclose(xi0, g["xi0"][b], rtol=0, atol=1e-12)
    rows = reach[g0] if standoff else goals[g0][None]
    return P.plan(robot, sc, cfg, xi0, g["start"][b], goals[g0], rows, goal_set=goals, reach_grasps=reach,
                  goal_idx=g0, learner=learner)


def test_plan_fixtures_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[5:-4] for p in GOLDEN])
def test_oracle_plan_matches_reference_planner(path):
    g = np.load(path)
    sc = S.make_scene(**eval(str(g["scene_args"])))
    assert abs(sc["sdf_grids"].astype(np.float64).sum() - float(g["sdf_checksum"])) < 1e-6
    robot = R.PandaRef(body_points=g["body_points"])
    keys, fkeys = [str(k) for k in g["info_keys"]], [str(k) for k in g["flag_keys"]]
    for b in range(g["xi0"].shape[0]):
        hist, infos, selected, final = run_oracle_plan(g, sc, b, robot)
        assert len(hist) == int(g["history_len"][b]) and len(infos) == int(g["info_len"][b])
        np.testing.assert_allclose(np.stack(hist), g["history"][b, :len(hist)], rtol=0, atol=1e-9)
        np.testing.assert_allclose(final, g["final"][b], rtol=0, atol=1e-9)
        assert selected == g["selected"][b, :int(g["selected_len"][b])].tolist()
        for k, info in enumerate(infos):
            for c, key in enumerate(keys):
                want = g["infos"][b, k, c]
                slack = g["tie_slack"][b, k] * (1 + 1e-9) if key in ("obs", "cost") else 0.0
                assert abs(float(info[key]) - want) <= 1e-9 * max(1.0, abs(want)) + slack, (key, b, k)
            for c, key in enumerate(fkeys):
                assert int(bool(info[key])) == int(g["flags"][b, k, c]), (key, b, k)
