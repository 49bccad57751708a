#!/usr/bin/env python
"""Extract the Panda kinematic constants and synthetic body points the hot path needs.

Runs ONLY in the build container (reads /root/reference, which does not travel to the
GPU box).  Outputs are plain numeric data committed under omg_planner_b200/data/:

  panda_constants.json  -- from ycb_render/robotPose/robot_p3.pkl (loaded at
                           ycb_render/robotPose/robot_pykdl.py:98-112): _pose_0, _tip2joint,
                           _joint_axis, center_offset, joint limits (order of
                           omg/core.py:155-162, i.e. 7 arm joints + 2 fingers).
  panda_body_points.json -- SURVEY.md section 8(d): p=15 vertices per link sampled (seed 0)
                           from bullet/models/panda/meshes/collision/*.obj, mapped into the
                           mesh-centre frame with inv(center_offset[j]) because the FK output
                           is post-multiplied by center_offset (robot_pykdl.py:204-205).
                           (The reference samples data/robots/link*.xyz unseeded,
                           omg/core.py:166-190; data/ is absent.)
"""
import json
import os
import pickle
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "omg_planner_b200", "data")


def main():
    with open(os.path.join(REF, "ycb_render/robotPose/robot_p3.pkl"), "rb") as fid:
        info = pickle.load(fid)
    names = list(info["_joint_name"])
    del names[-3]  # dummy hand joint (omg/core.py:156)
    limits = [[float(info["_joint_limits"][n][0]), float(info["_joint_limits"][n][1])] for n in names]
    const = {
        "source": "ycb_render/robotPose/robot_p3.pkl",
        "pose_0": np.array(info["_pose_0"], dtype=np.float64).tolist(),
        "tip2joint": np.array(info["_tip2joint"], dtype=np.float64).tolist(),
        "joint_axis": np.array(info["_joint_axis"], dtype=np.float64).tolist(),
        "joint_origin_true": np.array(info["_joint_origin"], dtype=np.float64).tolist(),
        "center_offset": np.array(info["center_offset"], dtype=np.float64).tolist(),
        "joint_names": names,
        "joint_limits": limits,
    }
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "panda_constants.json"), "w") as f:
        json.dump(const, f, indent=1)

    links = ["link1", "link2", "link3", "link4", "link5", "link6", "link7", "hand", "finger", "finger"]
    rng = np.random.RandomState(0)
    co = np.array(info["center_offset"], dtype=np.float64)
    pts_all = []
    for j, name in enumerate(links):
        verts = []
        with open(os.path.join(REF, "bullet/models/panda/meshes/collision", name + ".obj")) as f:
            for line in f:
                if line.startswith("v "):
                    verts.append([float(t) for t in line.split()[1:4]])
        verts = np.array(verts)
        sel = rng.choice(verts.shape[0], 15, replace=False)
        v = verts[sel]
        inv = np.linalg.inv(co[j])
        pts_all.append((v @ inv[:3, :3].T + inv[:3, 3]).tolist())
    with open(os.path.join(OUT, "panda_body_points.json"), "w") as f:
        json.dump({"source": "bullet/models/panda/meshes/collision/*.obj, seed 0, 15 per link",
                   "points": pts_all}, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
