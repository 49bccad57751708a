#!/usr/bin/env python
"""Stage the handful of UNMODIFIED reference files that tests/test_gpu_reference_classes.py imports on the GPU box into
tests/_ref_snapshot/ (git-ignored: the files never enter this repository's history; like the built *.so files they
travel to the GPU box with the gpurun snapshot).  BUILD-CONTAINER ONLY (needs /root/reference).

The test proves the drop-in boundary with the reference's own classes on the B200:
  (i)  omg/cost.py + omg/optimizer.py + layers/sdf_matching_loss.py + robot_pykdl.py's FK, unmodified, over this
       repo's omg_cuda.sdf_loss_forward (INTEGRATION.md level 1);
  (ii) omg/planner.py's Planner.plan, unmodified, with `from .optimizer import Optimizer`, `from .cost import Cost`,
       `from .online_learner import Learner` resolving to this repo's classes (INTEGRATION.md level 2)."""
import filecmp
import os
import shutil
import sys

REF = os.environ.get("OMG_REFERENCE_ROOT_SRC", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "tests", "_ref_snapshot")
FILES = [
    "omg/__init__.py", "omg/cost.py", "omg/optimizer.py", "omg/planner.py", "omg/online_learner.py", "omg/config.py",
    "omg/util.py", "layers/__init__.py", "layers/sdf_matching_loss.py", "ycb_render/__init__.py",
    "ycb_render/robotPose/__init__.py", "ycb_render/robotPose/_init_paths.py", "ycb_render/robotPose/robot_pykdl.py",
    "ycb_render/robotPose/robot_p3.pkl", "LICENSE",
]


def stage(verbose=True):
    if not os.path.isdir(os.path.join(REF, "omg")):
        if verbose:
            print("no reference tree at", REF, "- nothing staged")
        return None
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified files of liruiw/OMG-Planner staged by tools/stage_ref_snapshot.py for the GPU-box tests.\n"
                "Git-ignored; not part of this repository.\n")
    if verbose:
        print("staged", len(FILES), "reference files under", DST)
    return DST


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
