"""`import omg_cuda` resolves here when the repo root is on sys.path: the reference's
layers/sdf_matching_loss.py:5 then binds the B200 operator without modification."""
from omg_planner_b200.omg_cuda import sdf_loss_forward  # noqa: F401
