"""Import the UNMODIFIED reference Python (omg.cost / omg.optimizer / omg.config / omg.util / the FK
method of ycb_render/robotPose/robot_pykdl.py) from /root/reference under stubs, with `omg_cuda`
bound to the CPU restatement of layers/sdf_matching_loss_kernel.cu (oracle/sdf_loss_ref.c).

It exists to (1) pin oracle/chomp_ref.py against the reference's own code, (2) generate tests/golden/*.npz
(tools/make_golden.py) -- both BUILD-CONTAINER ONLY, /root/reference does not exist on the GPU box -- and (3) drive
the reference's own classes on the B200 from the staged copy tests/_ref_snapshot/ (tools/stage_ref_snapshot.py;
OMG_REFERENCE_ROOT points at it) with `omg_cuda` bound to the product operator
(tests/test_gpu_reference_classes.py).  bench.py and __graft_entry__ never import this file.

Stub list follows SURVEY.md section 8(c).
"""
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("OMG_REFERENCE_ROOT", "/root/reference")
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _AttrDict(dict):
    """easydict.EasyDict stand-in: a dict with attribute access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _quat2mat(q):
    w, x, y, z = [float(t) for t in q]
    n = w * w + x * x + y * y + z * z
    s = 2.0 / n if n > 0 else 0.0
    return np.array(
        [
            [1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
            [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
            [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)],
        ]
    )


def _mat2quat(M):
    from scipy.spatial.transform import Rotation

    x, y, z, w = Rotation.from_matrix(np.asarray(M)).as_quat()
    q = np.array([w, x, y, z])
    return q if w >= 0 else -q


def _euler2mat(ai, aj, ak, axes="sxyz"):
    from scipy.spatial.transform import Rotation

    return Rotation.from_euler("xyz", [ai, aj, ak]).as_matrix()


def _mat2euler(M, axes="sxyz"):
    from scipy.spatial.transform import Rotation

    return tuple(Rotation.from_matrix(np.asarray(M)).as_euler("xyz"))


def _axangle2mat(axis, angle):
    from scipy.spatial.transform import Rotation

    axis = np.asarray(axis, dtype=float)
    return Rotation.from_rotvec(axis / np.linalg.norm(axis) * angle).as_matrix()


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless placeholder class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


def install_stubs(sdf_forward):
    """Register stub modules. `sdf_forward` becomes omg_cuda.sdf_loss_forward."""
    import torch

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # omg/cost.py:136,218,331-335; omg/config.py:222-227
    if not hasattr(np, "int"):
        np.int = int  # removed alias used at omg/sdf_tools.py:48
    mods = {}
    mods["IPython"] = _Anything("IPython")
    ed = types.ModuleType("easydict")
    ed.EasyDict = _AttrDict
    mods["easydict"] = ed
    t3 = types.ModuleType("transforms3d")
    tq = types.ModuleType("transforms3d.quaternions")
    tq.quat2mat, tq.mat2quat = _quat2mat, _mat2quat
    te = types.ModuleType("transforms3d.euler")
    te.euler2mat, te.mat2euler = _euler2mat, _mat2euler
    ta = types.ModuleType("transforms3d.axangles")
    ta.axangle2mat = _axangle2mat
    t3.quaternions, t3.euler, t3.axangles = tq, te, ta
    mods.update({"transforms3d": t3, "transforms3d.quaternions": tq, "transforms3d.euler": te,
                 "transforms3d.axangles": ta})
    mods["PyKDL"] = _Anything("PyKDL")
    oc = types.ModuleType("omg_cuda")
    oc.sdf_loss_forward = sdf_forward
    mods["omg_cuda"] = oc
    for k, v in mods.items():
        sys.modules[k] = v
    # robot_pykdl.py:25-26 imports the URDF/KDL parsers (need lxml + PyKDL): placeholders.
    for name in ("ycb_render.robotPose.kdl_parser", "ycb_render.robotPose.urdf_parser_py",
                 "ycb_render.robotPose.urdf_parser_py.urdf"):
        sys.modules[name] = _Anything(name)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


_loaded = {}


def load_reference(sdf_forward=None):
    """Returns a namespace with the reference modules: cost, optimizer, config, util, robot_pykdl."""
    if _loaded:
        return _loaded["ns"]
    if sdf_forward is None:
        sys.path.insert(0, REPO_ROOT)
        from oracle import sdf_loss_ref

        sdf_forward = sdf_loss_ref.sdf_loss_forward_torch
    install_stubs(sdf_forward)
    ns = types.SimpleNamespace()
    ns.config = importlib.import_module("omg.config")
    ns.util = importlib.import_module("omg.util")
    ns.cost = importlib.import_module("omg.cost")
    ns.optimizer = importlib.import_module("omg.optimizer")
    ns.online_learner = importlib.import_module("omg.online_learner")
    ns.robot_pykdl = importlib.import_module("ycb_render.robotPose.robot_pykdl")
    ns.cfg = ns.config.cfg
    _loaded["ns"] = ns
    return ns


def make_ref_kinematics(ns):
    """robot_kinematics without __init__ (needs URDF + PyKDL): set exactly the fields the FK method
    reads (robot_pykdl.py:101-110), including the joint-origin aliasing of line 104."""
    import pickle

    rk = ns.robot_pykdl.robot_kinematics.__new__(ns.robot_pykdl.robot_kinematics)
    with open(os.path.join(REF_ROOT, "ycb_render/robotPose/robot_p3.pkl"), "rb") as fid:
        info = pickle.load(fid)
    rk._pose_0 = info["_pose_0"]
    rk._joint_origin = info["_joint_axis"]  # sic, robot_pykdl.py:104
    rk._tip2joint = info["_tip2joint"]
    rk._joint_axis = info["_joint_axis"]
    rk._joint_limits = info["_joint_limits"]
    rk._joint_name = info["_joint_name"]
    rk.center_offset = np.array(info["center_offset"])
    return rk


class RefTrajectory(object):
    """Same semantics as omg/core.py:23-57 (Trajectory.update / .set) without importing omg.core
    (which drags in the OpenGL renderer)."""

    def __init__(self, ns, data, start, end, goal_set=None, goal_idx=0):
        self.ns = ns
        self.data = np.array(data, dtype=np.float64)
        self.start = np.array(start, dtype=np.float64)
        self.end = np.array(end, dtype=np.float64)
        self.goal_set = goal_set if goal_set is not None else []
        self.goal_idx = goal_idx

    def update(self, grad):  # omg/core.py:43-51
        if self.ns.cfg.consider_finger:
            self.data += grad
        else:
            self.data[:, :-2] += grad[:, :-2]
        self.data[:, -2:] = np.minimum(np.maximum(self.data[:, -2:], 0), 0.04)

    def set(self, new_traj):  # omg/core.py:53-57
        self.data = new_traj

    def interpolate_waypoints(self, waypoints=None, mode="cubic"):  # omg/core.py:59-78 (dynamic_timestep off)
        self.data = self.ns.util.interpolate_waypoints(np.stack([self.start, self.end]), self.ns.cfg.timesteps,
                                                       self.start.shape[0], mode=mode)


def make_ref_env(ns, scene, body_points):
    """Stand-in for omg.core.Env carrying exactly what Cost/Optimizer read (SURVEY 8b 'Scene inputs')."""
    import torch

    env = types.SimpleNamespace()
    env.config = ns.cfg
    env.target_idx = int(scene["target_idx"])
    env.objects = []
    for i, name in enumerate(scene["names"]):
        o = types.SimpleNamespace()
        o.name = name
        o.pose_mat = np.array(scene["pose_mats"][i], dtype=np.float64)
        o.attached = False
        o.reach_grasps = []
        o.grasps = []
        env.objects.append(o)
    env.sdf_torch = torch.from_numpy(np.ascontiguousarray(scene["sdf_grids"], dtype=np.float32))
    env.sdf_limits = torch.from_numpy(np.ascontiguousarray(scene["sdf_limits"], dtype=np.float32))
    robot = types.SimpleNamespace()
    robot.robot_kinematics = make_ref_kinematics(ns)
    robot.collision_points = np.array(body_points, dtype=np.float64)
    names = list(robot.robot_kinematics._joint_name)
    del names[-3]
    lim = robot.robot_kinematics._joint_limits
    robot.joint_lower_limit = np.array([[lim[n][0] for n in names]])
    robot.joint_upper_limit = np.array([[lim[n][1] for n in names]])
    robot.joint_lower_limit[:, :-2] += ns.cfg.soft_joint_limit_padding  # omg/core.py:163-164
    robot.joint_upper_limit[:, :-2] -= ns.cfg.soft_joint_limit_padding
    env.robot = robot
    return env
