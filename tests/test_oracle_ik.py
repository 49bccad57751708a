"""CPU: the C restatement of the reference's KDL inverse kinematics (oracle/kdl_ik_ref.c) against (1) fixtures
recorded from the reference's own KDL sources compiled here (tests/golden/ik_kdl.npz, tools/make_golden_goalset.py) and
(2) that library itself when it is present (oracle/_ref/libkdl_ik.so: built in the build container, travels to the GPU
box).  Byte-for-byte: both run the same IEEE operations in the same order."""
import os

import numpy as np
import pytest

from oracle import kdl_ik_ref as K

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ik_kdl.npz")


def _chain(g):
    return K.PandaChain(g["pose_0"], g["lower"], g["upper"])


def test_restatement_reproduces_kdl_fixture_bit_for_bit():
    g = np.load(GOLD)
    ch = _chain(g)
    P, S = g["status"].shape
    assert (g["status"] >= 0).sum() > 50 and (g["status"] < 0).sum() > 50
    for p in range(P):
        for s in range(S):
            q, rc, steps, raw = ch.ik(g["targets"][p, :3], g["targets"][p, 3:], g["seeds"][s])
            assert rc == g["status"][p, s], (p, s)
            np.testing.assert_array_equal(raw, g["sols"][p, s])
            if rc >= 0:   # a solution is a solution: FK lands on the target, joints inside the limits
                T = ch.fk_hand(q)
                assert np.abs(T[:3, 3] - g["targets"][p, :3]).max() < 2e-6
                assert (q >= g["lower"] - 1e-12).all() and (q <= g["upper"] + 1e-12).all()
                assert steps < 100


@pytest.mark.skipif(not K.have_ref(), reason="oracle/_ref/libkdl_ik.so (the reference's KDL) is not built here")
def test_restatement_equals_reference_kdl_on_fresh_problems():
    g = np.load(GOLD)
    ch = _chain(g)
    rng = np.random.RandomState(77)
    solved = 0
    for k in range(150):
        q = rng.uniform(ch.lo, ch.hi)
        np.testing.assert_array_equal(ch.fk_hand(q), ch.ref_fk_hand(q))
        T = ch.fk_hand(q)
        from omg_planner_b200.ik import poses_to_targets
        tg = poses_to_targets(T)
        if k % 4 == 0:
            tg[:3] += rng.uniform(-0.3, 0.3, 3)
        seed = rng.uniform(ch.lo, ch.hi)
        a, rc_a, raw_a = ch.ref_ik(tg[:3], tg[3:], seed)
        b, rc_b, _, raw_b = ch.ik(tg[:3], tg[3:], seed)
        assert rc_a == rc_b
        np.testing.assert_array_equal(raw_a, raw_b)
        solved += rc_a >= 0
    assert 20 < solved < 150


def test_chain_of_solves_stops_at_first_failure():
    g = np.load(GOLD)
    ch = _chain(g)
    ok = np.argwhere(g["status"] >= 0)
    bad = np.argwhere(g["status"] < 0)
    p, s = ok[0]
    pb = bad[bad[:, 1] == s][0][0]
    n, sols = ch.ik_chain(np.stack([g["targets"][p], g["targets"][p]]), g["seeds"][s])
    assert n == 2
    np.testing.assert_array_equal(sols[0], g["sols"][p, s])
    n, _ = ch.ik_chain(np.stack([g["targets"][pb], g["targets"][p]]), g["seeds"][s])
    assert n == 0
