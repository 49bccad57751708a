"""CPU: the oracle (oracle/chomp_ref.py) replayed against fixtures produced by the reference's own
Python (tools/make_golden.py -> tests/golden/chomp_*.npz)."""
import glob
import os

import numpy as np
import pytest

import helpers as H
from omg_planner_b200 import scene as S
from oracle import chomp_ref as R

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "chomp_*.npz")))


def load_case(path):
    g = np.load(path)
    args = eval(str(g["scene_args"]))  # repr of a plain dict written by tools/make_golden.py
    sc = S.make_scene(**args)
    assert abs(sc["sdf_grids"].astype(np.float64).sum() - float(g["sdf_checksum"])) < 1e-6, "scene generator drifted"
    return g, sc, H.mode_from_fixture(g)


def test_fixtures_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[6:-4] for p in GOLDEN])
def test_oracle_matches_reference_history(path):
    g, sc, mode = load_case(path)
    robot = R.PandaRef(body_points=g["body_points"])
    hist, infos, flags = g["history"], g["infos"], g["flags"]
    keys, fkeys = [str(k) for k in g["info_keys"]], [str(k) for k in g["flag_keys"]]
    n_traj, n_iter = hist.shape[0], hist.shape[1] - 1
    for b in range(n_traj):
        cfg = R.RefConfig(**mode)
        rows = None
        if mode["goal_set_proj"]:
            rows = g["tails"][b] if mode["use_standoff"] else g["end"][b][None]
        opt = R.ChompRef(robot, sc, cfg, g["xi0"][b], g["start"][b], g["end"][b], rows)
        for it in range(n_iter):
            info = opt.step()
            np.testing.assert_allclose(opt.xi, hist[b, it + 1], rtol=0, atol=1e-10)
            np.testing.assert_allclose(info["gradient"], g["grads"][b, it], rtol=1e-9, atol=1e-9)
            for k, key in enumerate(keys):
                assert abs(float(info[key]) - infos[b, it, k]) <= 1e-9 * max(1.0, abs(infos[b, it, k])), (key, it)
            for k, key in enumerate(fkeys):
                assert int(bool(info[key])) == int(flags[b, it, k]), (key, it)
