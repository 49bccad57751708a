"""Panda constants for the hot path (host side).

Values come from ycb_render/robotPose/robot_p3.pkl (read at robot_pykdl.py:98-112) and the padded joint
limits of omg/core.py:157-164; extracted once by tools/extract_robot_constants.py into data/*.json."""
import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class PandaConstants(object):
    def __init__(self, body_points=None, soft_joint_limit_padding=0.2):
        with open(os.path.join(_DATA, "panda_constants.json")) as f:
            c = json.load(f)
        self.pose_0 = np.ascontiguousarray(c["pose_0"], dtype=np.float64)
        self.tip2joint = np.ascontiguousarray(c["tip2joint"], dtype=np.float64)
        self.joint_axis = np.ascontiguousarray(c["joint_axis"], dtype=np.float64)
        self.joint_origin_true = np.ascontiguousarray(c["joint_origin_true"], dtype=np.float64)
        self.center_offset = np.ascontiguousarray(c["center_offset"], dtype=np.float64)
        lim = np.array(c["joint_limits"], dtype=np.float64)
        self.joint_lower_limit = lim[None, :, 0].copy()
        self.joint_upper_limit = lim[None, :, 1].copy()
        self.joint_lower_limit[:, :-2] += soft_joint_limit_padding
        self.joint_upper_limit[:, :-2] -= soft_joint_limit_padding
        if body_points is None:
            with open(os.path.join(_DATA, "panda_body_points.json")) as f:
                body_points = json.load(f)["points"]
        self.collision_points = np.ascontiguousarray(body_points, dtype=np.float64)  # [10, p, 3]
