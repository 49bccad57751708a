"""ORACLE -- test infrastructure only.

CPU restatement of the reference's CHOMP hot path (omg/optimizer.py + omg/cost.py +
layers/sdf_matching_loss_kernel.cu + the FK method of ycb_render/robotPose/robot_pykdl.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (omg_planner_b200) never does.
"""
