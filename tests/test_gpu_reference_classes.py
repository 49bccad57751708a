"""GPU: the drop-in boundary proven with THE REFERENCE'S OWN CLASSES on the B200.

The unmodified reference files are imported from tests/_ref_snapshot/ (staged by tools/stage_ref_snapshot.py in the
build container; git-ignored, it travels to the GPU box like the built libraries) under the stub modules of
tools/ref_harness.py (IPython, easydict, transforms3d, PyKDL -- SURVEY.md 8c).

  (i)  INTEGRATION.md level 1: omg/cost.py:Cost + omg/optimizer.py:Optimizer + layers/sdf_matching_loss.py:SDFLoss +
       robot_pykdl.py's FK, unmodified, with `import omg_cuda` (layers/sdf_matching_loss.py:5) resolving to this repo's
       operator (omgb_sdf_loss on the B200): 25 iterations == tests/golden/chomp_*.npz.
  (ii) INTEGRATION.md level 2: omg/planner.py:Planner.plan, unmodified, with its imports `from .optimizer import
       Optimizer`, `from .cost import Cost`, `from .online_learner import Learner` (planner.py:5-8) resolving to this
       repo's plugin classes: == tests/golden/plan_*.npz (history, info list, selected goals)."""
import builtins
import glob
import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SNAP = os.path.join(HERE, "_ref_snapshot")

pytestmark = pytest.mark.gpu

CHOMP = sorted(glob.glob(os.path.join(HERE, "golden", "chomp_*.npz")))
PLANS = sorted(glob.glob(os.path.join(HERE, "golden", "plan_*.npz")))
TOL_RAD = 1e-7


@pytest.fixture(scope="module")
def ref():
    """The reference modules (from the staged snapshot) with omg_cuda = the product operator."""
    if not os.path.isfile(os.path.join(SNAP, "omg", "cost.py")):
        try:   # build container: stage on the fly
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import stage_ref_snapshot
            stage_ref_snapshot.stage(verbose=False)
        except Exception:   # noqa: BLE001
            pass
    if not os.path.isfile(os.path.join(SNAP, "omg", "cost.py")):
        pytest.skip("tests/_ref_snapshot/ not staged (run tools/stage_ref_snapshot.py in the build container)")
    os.environ["OMG_REFERENCE_ROOT"] = SNAP
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_harness as RH
    assert RH.REF_ROOT == SNAP, "tools/ref_harness.py was imported before with another reference root"
    from omg_planner_b200.engine import sdf_loss_forward
    ns = RH.load_reference(sdf_forward=sdf_loss_forward)
    assert os.path.dirname(os.path.abspath(ns.cost.__file__)) == os.path.join(SNAP, "omg")
    import omg_cuda as bound   # what layers/sdf_matching_loss.py:5 imported
    assert bound.sdf_loss_forward is sdf_loss_forward
    ns.RH = RH
    ns.cfg.timeout = -1
    ns.cfg.report_cost = False
    ns.cfg.report_time = False
    ns.cfg.silent = True
    return ns


def _env(ns, sc, body_points):
    env = ns.RH.make_ref_env(ns, sc, body_points)
    env.sdf_torch = env.sdf_torch.cuda()      # what Env.combine_sdfs hands over (omg/core.py:372-411)
    env.sdf_limits = env.sdf_limits.cuda()
    return env


@pytest.mark.parametrize("path", CHOMP, ids=[os.path.basename(p)[6:-4] for p in CHOMP])
def test_reference_cost_and_optimizer_over_the_b200_operator(ref, path):
    import helpers as H
    from omg_planner_b200 import _lib
    from omg_planner_b200 import scene as S

    ns, cfg = ref, ref.cfg
    g = np.load(path)
    mode = H.mode_from_fixture(g)
    cfg.consider_finger = False
    for k, v in mode.items():
        cfg[k] = v
    cfg.timesteps = g["xi0"].shape[1]
    ns.config.get_global_param(cfg.timesteps)
    sc = S.make_scene(**eval(str(g["scene_args"])))
    assert float(sc["sdf_grids"].astype(np.float64).sum()) == float(g["sdf_checksum"])
    env = _env(ns, sc, g["body_points"])
    keys, fkeys = [str(k) for k in g["info_keys"]], [str(k) for k in g["flag_keys"]]
    launches0 = int(_lib.lib().omgb_launch_count())
    worst = 0.0
    B, iters = g["xi0"].shape[0], g["history"].shape[1] - 1
    for b in range(B):
        cost = ns.cost.Cost(env)
        opt = ns.optimizer.Optimizer(env, cost)
        traj = ns.RH.RefTrajectory(ns, g["xi0"][b], g["start"][b], g["end"][b], goal_set=[g["end"][b]], goal_idx=0)
        if cfg.goal_set_proj:
            env.objects[env.target_idx].reach_grasps = [g["tails"][b]] if cfg.use_standoff else [g["end"][b]]
            cost.target_obj = env.objects[env.target_idx]
        for it in range(iters):
            info = opt.optimize(traj, force_update=True)
            err = np.abs(traj.data - g["history"][b, it + 1])[:, :7].max()
            worst = max(worst, err)
            assert err <= TOL_RAD, (b, it, err)
            for c, key in enumerate(keys):
                want = g["infos"][b, it, c]
                slack = g["tie_slack"][b, it] * (1 + 1e-9) if key in ("obs", "cost") else 0.0
                assert abs(float(info[key]) - want) <= 1e-6 * max(1.0, abs(want)) + slack, (key, b, it)
            for c, key in enumerate(fkeys):
                assert int(bool(info[key])) == int(g["flags"][b, it, c]), (key, b, it)
    # every operator call of the reference's Cost went through libomgb200.so (2 launches per omgb_sdf_loss call)
    assert int(_lib.lib().omgb_launch_count()) - launches0 >= 2 * B * iters
    print("reference Cost/Optimizer over omgb_sdf_loss: worst |xi - fixture| = %.2e rad" % worst)


def _planner_module_over_plugin_classes(ns):
    """omg/planner.py imported UNMODIFIED with `.optimizer`, `.cost`, `.online_learner` resolving to this repo's modules
    (the import swap of INTEGRATION.md section 2, done through sys.modules instead of editing the file)."""
    from omg_planner_b200 import cost as our_cost
    from omg_planner_b200 import online_learner as our_learner
    from omg_planner_b200 import optimizer as our_optimizer

    names = {"omg.optimizer": our_optimizer, "omg.cost": our_cost, "omg.online_learner": our_learner}
    saved = {k: sys.modules.get(k) for k in list(names) + ["omg.planner"]}
    try:
        sys.modules.update(names)
        sys.modules.pop("omg.planner", None)
        mod = importlib.import_module("omg.planner")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert os.path.abspath(mod.__file__) == os.path.join(SNAP, "omg", "planner.py")
    assert mod.Optimizer is our_optimizer.Optimizer and mod.Cost is our_cost.Cost and mod.Learner is our_learner.Learner
    return mod


@pytest.mark.parametrize("path", PLANS, ids=[os.path.basename(p)[5:-4] for p in PLANS])
def test_reference_planner_plan_over_the_plugin_classes(ref, path):
    from omg_planner_b200 import scene as S

    ns, cfg = ref, ref.cfg
    pm = _planner_module_over_plugin_classes(ns)
    g = np.load(path)
    cfg.goal_set_proj, cfg.use_standoff = bool(g["goal_set_proj"]), bool(g["use_standoff"])
    cfg.ol_alg = str(g["ol_alg"])
    cfg.top_k_collision, cfg.consider_finger = 1000, False
    cfg.pre_terminate = bool(g["pre_terminate"])
    cfg.optim_steps, cfg.extra_smooth_steps = int(g["optim_steps"]), int(g["extra_smooth_steps"])
    n = g["xi0"].shape[1]
    cfg.timesteps = n
    ns.config.get_global_param(n)
    sc = S.make_scene(**eval(str(g["scene_args"])))
    env = _env(ns, sc, g["body_points"])
    keys, fkeys = [str(k) for k in g["info_keys"]], [str(k) for k in g["flag_keys"]]
    learner_on = cfg.goal_set_proj and cfg.ol_alg not in ("Baseline", "Proj")
    worst = 0.0
    for b in range(g["xi0"].shape[0]):
        target = env.objects[env.target_idx]
        cost = pm.Cost(env)
        optim = pm.Optimizer(env, cost)
        if learner_on:
            target.reach_grasps = g["reach"][b] if cfg.use_standoff else g["goals"][b]
            traj = ns.RH.RefTrajectory(ns, np.zeros((n, 9)), g["start"][b], g["goals"][b, 0], goal_set=list(g["goals"][b]),
                                       goal_idx=0)
            traj.interpolate_waypoints()
        else:
            traj = ns.RH.RefTrajectory(ns, g["xi0"][b], g["start"][b], g["end"][b], goal_set=[g["end"][b]], goal_idx=0)
            if cfg.goal_set_proj:
                target.reach_grasps = [g["tails"][b]]
        cost.target_obj = target
        p = pm.Planner.__new__(pm.Planner)   # (__init__ loads grasp files and runs IK; plan() is the reference's)
        p.cfg, p.env, p.traj, p.cost, p.optim = cfg, env, traj, cost, optim
        if learner_on:
            p.learner = pm.Learner(env, traj, cost)
        assert np.abs(traj.data - g["xi0"][b]).max() <= 1e-12
        _print = builtins.print
        builtins.print = lambda *a, **k: None
        try:
            info = p.plan(traj)
        finally:
            builtins.print = _print
        hist = p.history_trajectories
        assert len(hist) == int(g["history_len"][b]) and len(info) == int(g["info_len"][b])
        err = np.abs(np.stack(hist) - g["history"][b, :len(hist)])[..., :7].max()
        worst = max(worst, err)
        assert err <= TOL_RAD, (b, err)
        assert np.abs(traj.data - g["final"][b])[..., :7].max() <= TOL_RAD
        assert [int(s) for s in p.selected_goals] == g["selected"][b, :int(g["selected_len"][b])].tolist()
        for k, i in enumerate(info):
            for c, key in enumerate(keys):
                want = g["infos"][b, k, c]
                slack = g["tie_slack"][b, k] * (1 + 1e-9) if key in ("obs", "cost") else 0.0
                assert abs(float(i[key]) - want) <= 1e-6 * max(1.0, abs(want)) + slack, (key, b, k)
            for c, key in enumerate(fkeys):
                assert int(bool(i[key])) == int(g["flags"][b, k, c]), (key, b, k)
        assert "time" in info[-1]
    print("reference Planner.plan over the plugin classes: worst |history - fixture| = %.2e rad" % worst)
