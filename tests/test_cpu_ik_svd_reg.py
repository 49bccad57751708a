"""CPU: the register-resident SVD of the IK kernel (omg_planner_b200/csrc/ik_svd_reg.cuh) compiled for the host and
checked bit for bit against the oracle's restatement of KDL's SVD_HH (tests/host/svd_reg_check.cpp): Jacobians of random
arm configurations, rank-deficient, badly scaled and zero-column matrices."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_register_svd_equals_oracle_svd_bit_for_bit(tmp_path):
    exe = str(tmp_path / "svd_reg_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++14", "-o", exe,
                           os.path.join(HERE, "host", "svd_reg_check.cpp"), "-lm"])
    out = subprocess.run([exe, "30000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    fields = dict(zip(out.stdout.split()[0::2], out.stdout.split()[1::2]))
    assert int(fields["mismatches"]) == 0 and int(fields["trials"]) == 30000
    assert int(fields["cancellation_branches"]) > 1000      # the split / cancellation path is exercised
