"""GPU: BASELINE config 2 sizes (1024 trajectories x 30 waypoints, 10 SDFs at 128^3) through
size-independent properties, plus an oracle spot check on a few trajectories of the big batch."""
import numpy as np
import pytest
import torch

import helpers as H
from omg_planner_b200 import scene as S
from omg_planner_b200.config import ChompConfig
from omg_planner_b200.robot import PandaConstants

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.fixture(scope="module")
def big():
    sc = S.make_scene(num_objects=10, grid=128, seed=0)
    robot = PandaConstants()
    xi, st, en, tails = S.make_trajectories(1024, 30, robot.joint_lower_limit, robot.joint_upper_limit, seed=0)
    return sc, robot, xi, st, en, tails


@pytest.mark.parametrize("name", ["goalset_standoff_topk", "fixed_full"])
def test_batch_properties(big, name):
    sc, robot, xi, st, en, tails = big
    mode = H.MODES[name]
    cfg = ChompConfig(**mode)
    cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(1)   # iteration 1 of a plan
    eng = H.engine_for(sc, cfg, robot)
    rows = H.goal_rows_for(mode, tails, en)
    r = None if rows is None else _dev(rows)
    # (a) determinism: the same step twice from the same state is bit-identical
    x1, x2 = _dev(xi), _dev(xi)
    i1 = eng.step(cfg, x1, _dev(st), _dev(en), r)["info"]
    i2 = eng.step(cfg, x2, _dev(st), _dev(en), r)["info"]
    torch.cuda.synchronize()
    assert torch.equal(x1, x2) and torch.equal(i1, i2)
    # (b) batch independence: a trajectory's result does not depend on its neighbours in the batch
    sel = np.array([0, 17, 511, 1023])
    xs = _dev(xi[sel])
    eng.step(cfg, xs, _dev(st[sel]), _dev(en[sel]), None if rows is None else _dev(rows[sel]))
    torch.cuda.synchronize()
    assert torch.equal(xs, x1[torch.from_numpy(sel).cuda()])
    # (c) active mask: masked-out trajectories are untouched
    act = torch.zeros(1024, dtype=torch.uint8, device="cuda"); act[::2] = 1
    x3 = _dev(xi)
    eng.step(cfg, x3, _dev(st), _dev(en), r, active=act)
    torch.cuda.synchronize()
    assert torch.equal(x3[1::2], _dev(xi)[1::2]) and torch.equal(x3[::2], x1[::2])
    # (d) invariants of Trajectory.update + handle_joint_limit: fingers clamped, arm inside padded limits
    out = x1.cpu().numpy()
    assert (out[..., 7:] >= 0).all() and (out[..., 7:] <= 0.04).all()
    lo, hi = robot.joint_lower_limit[0, :7], robot.joint_upper_limit[0, :7]
    viol = np.maximum(lo - out[..., :7], 0) + np.maximum(out[..., :7] - hi, 0)
    rounds = i1[:, 14].cpu().numpy()
    assert (np.linalg.norm(viol.reshape(1024, -1), axis=1)[rounds < 10] <= 1e-2 + 1e-12).all()
    if mode["goal_set_proj"]:   # the projected tail lands on the goal rows (up to the limit projection)
        tail_err = np.abs(out[:, -5:, :7] - tails[:, :, :7]).max(axis=(1, 2))
        assert (tail_err[rounds == 0] < 1e-9).all()
    # (e) oracle spot check on members of the big batch
    ref_hist, ref_infos = H.oracle_steps(sc, mode, xi[sel], st[sel], en[sel], None if rows is None else rows[sel], 1)
    err = np.abs(out[sel] - ref_hist[:, 1])[..., :7].max()
    print(name, "spot-check max |dxi|:", err, "P_in mean:", i1[:, 12].mean().item(), "nnz mean:", i1[:, 13].mean().item())
    assert err <= 1e-7
    for k, b in enumerate(sel):
        assert abs(i1[b, 0].item() - ref_infos[k][0]["obs"]) <= (1e-6 * max(1.0, abs(ref_infos[k][0]["obs"]))
                                                                   + ref_infos[k][0]["tie_slack"] * (1 + 1e-9))
        assert i1[b, 3].item() == ref_infos[k][0]["collide"]


def test_long_trajectories_and_many_objects():
    """60 waypoints x 20 objects (BASELINE config 4 shape, reduced grid) and 50 x 30 (config 5 shape)."""
    for n, o in ((60, 20), (50, 30)):
        sc = S.make_scene(num_objects=o, grid=32, seed=n)
        robot = PandaConstants()
        mode = H.MODES["goalset_standoff_topk"]
        cfg = ChompConfig(timesteps=n, **mode)
        xi, st, en, tails = S.make_trajectories(6, n, robot.joint_lower_limit, robot.joint_upper_limit, seed=n)
        ref_hist, _ = H.oracle_steps(sc, mode, xi, st, en, tails, 2)
        eng = H.engine_for(sc, cfg, robot)
        x = _dev(xi)
        for it in range(2):
            cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
            eng.step(cfg, x, _dev(st), _dev(en), _dev(tails))
        torch.cuda.synchronize()
        assert np.abs(x.cpu().numpy() - ref_hist[:, 2])[..., :7].max() <= 1e-7


def test_exact_accelerations_do_not_change_results(big):
    """Lower-bound culling and longest-first CTA order are exact: outputs (including P_in) are bit-identical with
    them off."""
    sc, robot, xi, st, en, tails = big
    outs = []
    for lb, lpt in ((1, 1), (0, 0)):
        mode = H.MODES["goalset_standoff_topk"]
        cfg = ChompConfig(**mode)
        eng = H.engine_for(sc, cfg, robot)
        eng.set_options(use_lower_bound=lb, use_longest_first=lpt)
        x = _dev(xi[:400])
        infos = []
        for it in range(3):
            cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
            infos.append(eng.step(cfg, x, _dev(st[:400]), _dev(en[:400]), _dev(tails[:400]))["info"][:, :15].clone())
        torch.cuda.synchronize()
        outs.append((x.clone(), torch.stack(infos)))
    assert torch.equal(outs[0][0], outs[1][0])                       # trajectories: bit-identical
    a, b = outs[0][1], outs[1][1]
    assert torch.equal(a[..., 3:], b[..., 3:]) and torch.equal(a[..., 1], b[..., 1])   # collide, P_in, flags, norms
    # obstacle cost / total cost are block sums whose fp64 summation order follows the active set: last-ulp only
    assert torch.allclose(a[..., [0, 2]], b[..., [0, 2]], rtol=1e-12, atol=0)


def test_persistent_plan_kernel_equals_per_iteration_launches(big):
    """omgb_chomp_plan runs the whole plan as one persistent launch with a device-side work queue; every trajectory must
    end bit-identical to the same number of per-iteration launches (also with stop_on_terminate freezing)."""
    sc, robot, xi, st, en, tails = big
    mode = H.MODES["goalset_standoff_topk"]
    cfg = ChompConfig(**mode)
    eng = H.engine_for(sc, cfg, robot)
    iters = 9
    x1 = _dev(xi)
    for it in range(iters):
        cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
        info1 = eng.step(cfg, x1, _dev(st), _dev(en), _dev(tails))["info"]
    for rep in range(2):   # twice: the queue state is reset per call
        x2 = _dev(xi)
        out = eng.plan(cfg, x2, _dev(st), _dev(en), _dev(tails), iters=iters)
        torch.cuda.synchronize()
        assert torch.equal(x1, x2)
        assert torch.equal(info1[:, :15], out["info"][:, :15])
    # more trajectories than resident CTA slots and a batch that is not a multiple of anything
    sel = slice(0, 1001)
    x3 = _dev(xi[sel])
    eng.plan(cfg, x3, _dev(st[sel]), _dev(en[sel]), _dev(tails[sel]), iters=iters)
    torch.cuda.synchronize()
    assert torch.equal(x3, x1[sel])


@pytest.mark.parametrize("name", ["goalset_standoff_topk", "fixed_full"])
def test_bricked_quad_layout_is_bit_identical(big, name):
    """omgb_scene_set_sdf_layout(1): the exact path reads the bricked quad copy of the grids (two 128-bit loads per
    trilinear sample) instead of the reference [O,X,Y,Z] layout; same taps, same lerps -> identical bits, in the
    per-point phase (value only), the winners (7 samples) and the full-sum path, and in the goal-scoring kernel."""
    sc, robot, xi, st, en, tails = big
    mode = H.MODES[name]
    outs = []
    for layout in ("plain", "quad"):
        cfg = ChompConfig(**mode)
        from omg_planner_b200.engine import ChompEngine
        eng = ChompEngine(robot=robot).load_scene(sc, cfg, sdf_layout=layout)
        x = _dev(xi[:300])
        rows = _dev(H.goal_rows_for(mode, tails[:300], en[:300])) if mode["goal_set_proj"] else None
        infos = []
        for it in range(4):
            cfg.obstacle_weight, cfg.smoothness_weight, cfg.step_size = cfg.schedule(it + 1)
            infos.append(eng.step(cfg, x, _dev(st[:300]), _dev(en[:300]), rows, want_grad=True))
        goals = _dev(en[:7])
        gc = eng.goal_costs(_dev(xi[:64]), 3, goals)
        torch.cuda.synchronize()
        outs.append((x.clone(), torch.stack([i["info"] for i in infos]), infos[-1]["grad"].clone(), gc.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    assert float(outs[0][1][..., 13].sum()) > 0    # non-zero potentials were evaluated
