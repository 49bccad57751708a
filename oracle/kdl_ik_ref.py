"""TEST INFRASTRUCTURE -- ctypes bindings of (1) oracle/kdl_ik_ref.c, the C restatement of the reference's KDL inverse
kinematics, and (2) oracle/_ref/libkdl_ik.so, the reference's OWN vendored KDL compiled from /root/reference by
oracle/kdl_ref/Makefile (build container only; the .so travels to the GPU box).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libomg_oracle_ik.so")
_REF_SO = os.path.join(_HERE, "_ref", "libkdl_ik.so")
_vp = ctypes.c_void_p
_lib = None
_ref = None


def build(force=False):
    src = os.path.join(_HERE, "kdl_ik_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def build_ref(reference_root="/root/reference"):
    """Compile the reference's KDL (only where /root/reference exists).  Returns the path or None."""
    if not os.path.isdir(os.path.join(reference_root, "orocos_kinematics_dynamics")):
        return _REF_SO if os.path.exists(_REF_SO) else None
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "kdl_ref"), "REF=" + reference_root])
    return _REF_SO


def have_ref():
    return os.path.exists(_REF_SO)


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.omg_oracle_ik.argtypes = [_vp] * 8
        _lib.omg_oracle_ik_chain.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp]
        _lib.omg_oracle_fk_hand.argtypes = [_vp] * 3
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class PandaChain(object):
    """frames = robot_kinematics._pose_0[:8] (7 joints + the fixed hand), limits = the padded arm limits
    (robot_pykdl.py:123-138)."""

    def __init__(self, pose_0, lower, upper):
        self.frames = _c(np.asarray(pose_0)[:8])
        self.lo = _c(np.asarray(lower).reshape(-1)[:7])
        self.hi = _c(np.asarray(upper).reshape(-1)[:7])

    # ---- the C restatement -----------------------------------------------------------------------------
    def ik(self, position, quat_xyzw, seed):
        """robot_kinematics.inverse_kinematics: (solution [7] or None, status, Newton steps)."""
        L = _load()
        pos, quat, seed = _c(position), _c(quat_xyzw), _c(np.asarray(seed)[:7])
        out = np.zeros(7)
        its = ctypes.c_int(0)
        rc = L.omg_oracle_ik(self.frames.ctypes.data, self.lo.ctypes.data, self.hi.ctypes.data, pos.ctypes.data,
                             quat.ctypes.data, seed.ctypes.data, out.ctypes.data, ctypes.addressof(its))
        return (out if rc >= 0 else None), rc, its.value, out

    def ik_chain(self, targets, seed):
        """targets [T,7] (position, quaternion xyzw), each solve seeded with the previous solution; returns
        (number solved, sols [T,7])."""
        L = _load()
        t, seed = _c(targets), _c(np.asarray(seed)[:7])
        sols = np.zeros((t.shape[0], 7))
        n = L.omg_oracle_ik_chain(self.frames.ctypes.data, self.lo.ctypes.data, self.hi.ctypes.data, t.ctypes.data,
                                  t.shape[0], seed.ctypes.data, sols.ctypes.data)
        return n, sols

    def fk_hand(self, q):
        L = _load()
        q = _c(np.asarray(q)[:7])
        out = np.zeros(16)
        L.omg_oracle_fk_hand(self.frames.ctypes.data, q.ctypes.data, out.ctypes.data)
        return out.reshape(4, 4)

    # ---- the reference's own KDL ---------------------------------------------------------------------------
    def _ref_handle(self):
        global _ref
        if _ref is None:
            _ref = ctypes.CDLL(_REF_SO)
            _ref.kdl_ik_create.restype = _vp
            _ref.kdl_ik_create.argtypes = [_vp, _vp, _vp, ctypes.c_int, _vp, _vp]
            _ref.kdl_ik_solve.argtypes = [_vp] * 5
            _ref.kdl_fk.argtypes = [_vp] * 3
        if not hasattr(self, "_h"):
            axes = _c(np.tile([0.0, 0.0, 1.0], (8, 1)))            # URDF: axis xyz="0 0 1" on every arm joint
            mov = np.array([1] * 7 + [0], np.int32)
            self._h = _ref.kdl_ik_create(self.frames.ctypes.data, axes.ctypes.data, mov.ctypes.data, 8,
                                         self.lo.ctypes.data, self.hi.ctypes.data)
        return self._h

    def ref_ik(self, position, quat_xyzw, seed):
        h = self._ref_handle()
        pos, quat, seed = _c(position), _c(quat_xyzw), _c(np.asarray(seed)[:7])
        out = np.zeros(7)
        rc = _ref.kdl_ik_solve(h, pos.ctypes.data, quat.ctypes.data, seed.ctypes.data, out.ctypes.data)
        return (out if rc >= 0 else None), rc, out

    def ref_fk_hand(self, q):
        h = self._ref_handle()
        q = _c(np.asarray(q)[:7])
        out = np.zeros(16)
        _ref.kdl_fk(h, q.ctypes.data, out.ctypes.data)
        return out.reshape(4, 4)
