"""Batched inverse kinematics over libomgb200.so (omgb_ik_solve / omgb_hand_poses): the device replacement for
robot_kinematics.inverse_kinematics (ycb_render/robotPose/robot_pykdl.py:257-289, KDL ChainIkSolverPos_NR_JL) and for
the hand-frame slice of forward_kinematics_parallel used by the goal-set filters (omg/planner.py:262-283).
There is no CPU fallback."""
import ctypes

import numpy as np
import torch

from . import _lib

_vp = ctypes.c_void_p


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def mat2quat_xyzw(R):
    """Unit quaternion (x, y, z, w), w >= 0, of rotation matrices [...,3,3] (what util.pack_pose + ros_quat hand to
    PyKDL: omg/util.py:105-127, 223-227).  KDL rebuilds the matrix from it with a formula that is even in q
    (frames.cpp:191-198), so only the rotation matters, not the sign convention."""
    R = np.asarray(R, dtype=np.float64)
    m = R.reshape(-1, 3, 3)
    # Bar-Itzhack: eigenvector of the symmetric 4x4 K matrix for the largest eigenvalue (robust to slightly
    # non-orthonormal input); all poses in one stacked eigh
    K = np.zeros((m.shape[0], 4, 4))
    K[:, 0, 0] = m[:, 0, 0] - m[:, 1, 1] - m[:, 2, 2]
    K[:, 1, 0] = m[:, 0, 1] + m[:, 1, 0]
    K[:, 1, 1] = m[:, 1, 1] - m[:, 0, 0] - m[:, 2, 2]
    K[:, 2, 0] = m[:, 0, 2] + m[:, 2, 0]
    K[:, 2, 1] = m[:, 1, 2] + m[:, 2, 1]
    K[:, 2, 2] = m[:, 2, 2] - m[:, 0, 0] - m[:, 1, 1]
    K[:, 3, 0] = m[:, 2, 1] - m[:, 1, 2]
    K[:, 3, 1] = m[:, 0, 2] - m[:, 2, 0]
    K[:, 3, 2] = m[:, 1, 0] - m[:, 0, 1]
    K[:, 3, 3] = m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    vals, vecs = np.linalg.eigh(K / 3.0)                 # (lower triangle is what eigh reads)
    q = vecs[np.arange(m.shape[0]), :, np.argmax(vals, axis=1)]   # x, y, z, w
    q = np.where(q[:, 3:4] >= 0, q, -q)
    return q.reshape(R.shape[:-2] + (4,))


def poses_to_targets(poses):
    """[...,4,4] homogeneous poses -> [...,7] (position, quaternion xyzw)."""
    poses = np.asarray(poses, dtype=np.float64)
    return np.concatenate([poses[..., :3, 3], mat2quat_xyzw(poses[..., :3, :3])], axis=-1)


class IkSolver(object):
    """chain: robot_kinematics._pose_0[:8] (7 arm joints + the fixed hand); limits: the padded arm limits."""

    def __init__(self, pose_0, lower, upper, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("IkSolver needs a CUDA device; there is no CPU fallback")
        self.L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.frames = _c(np.asarray(pose_0)[:8])
        self.lo = _c(np.asarray(lower).reshape(-1)[:7])
        self.hi = _c(np.asarray(upper).reshape(-1)[:7])

    def solve_chains(self, targets, seeds, want_steps=False):
        """targets [P,T,7] (position, quaternion xyzw), seeds [S,7] -> (sols [P,S,T,7], solved [P,S](, steps))
        as numpy arrays; one launch."""
        t = torch.from_numpy(_c(targets)).to(self.device)
        s = torch.from_numpy(_c(np.asarray(seeds)[:, :7])).to(self.device)
        if t.dim() != 3 or t.shape[2] != 7:
            raise RuntimeError("targets must be [P,T,7]")
        P, T, S = t.shape[0], t.shape[1], s.shape[0]
        sols = torch.zeros((P, S, T, 7), dtype=torch.float64, device=self.device)
        solved = torch.zeros((P, S), dtype=torch.int32, device=self.device)
        steps = torch.zeros((P, S, T), dtype=torch.int32, device=self.device) if want_steps else None
        with torch.cuda.device(self.device):
            _lib.check(self.L.omgb_ik_solve(self.frames.ctypes.data, self.lo.ctypes.data, self.hi.ctypes.data,
                                            _vp(t.data_ptr()), P, T, _vp(s.data_ptr()), S, _vp(sols.data_ptr()),
                                            _vp(solved.data_ptr()), _vp(steps.data_ptr()) if want_steps else None,
                                            _vp(torch.cuda.current_stream().cuda_stream)), "omgb_ik_solve")
        out = (sols.cpu().numpy(), solved.cpu().numpy())
        return out + (steps.cpu().numpy(),) if want_steps else out

    def inverse_kinematics(self, position, orientation=None, seed=None):
        """robot_kinematics.inverse_kinematics for one pose: orientation is a quaternion xyzw; returns [7] or None."""
        if orientation is None:
            raise RuntimeError("position-only IK is not used on the goal-set path")
        seed = np.zeros(7) if seed is None else np.asarray(seed, dtype=np.float64)[:7]
        tgt = np.concatenate([np.asarray(position, dtype=np.float64), np.asarray(orientation, dtype=np.float64)])
        sols, solved = self.solve_chains(tgt[None, None], seed[None])
        return sols[0, 0, 0] if solved[0, 0] == 1 else None

    def hand_poses(self, joints):
        """[M,>=7] joint vectors -> [M,4,4] hand frames (forward_kinematics_parallel(...)[:, 7] without the
        degree round trip)."""
        q = torch.from_numpy(_c(np.asarray(joints).reshape(-1, np.asarray(joints).shape[-1]))).to(self.device)
        M = q.shape[0]
        out = torch.empty((M, 4, 4), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.omgb_hand_poses(self.frames.ctypes.data, _vp(q.data_ptr()), q.shape[1], M,
                                              _vp(out.data_ptr()), _vp(torch.cuda.current_stream().cuda_stream)),
                       "omgb_hand_poses")
        return out.cpu().numpy()
