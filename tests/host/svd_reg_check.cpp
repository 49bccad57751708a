// Host check of omg_planner_b200/csrc/ik_svd_reg.cuh (the register-resident SVD of the IK kernel, compiled here for
// the CPU) against the oracle's restatement of KDL's SVD_HH (oracle/kdl_ik_ref.c, itself bit-identical to the
// reference's compiled KDL): U, w, V and the status must agree bit for bit on Jacobians of random arm configurations,
// on rank-deficient and on badly scaled matrices.  Built and run by tests/test_cpu_ik_svd_reg.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>

extern "C" {
#include "../../oracle/kdl_ik_ref.c"
}
#define OMGB_SVD_TRACE 1
#include "../../omg_planner_b200/csrc/ik_svd_reg.cuh"

static unsigned long long rng_state = 88172645463325252ULL;
static double rnd() {   // xorshift, uniform in [-1, 1)
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

int main(int argc, char **argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 20000;
    // the Panda chain of the product's constants is not needed: any chain-like Jacobians do; use random frames
    double frames[8 * 16];
    for (int s = 0; s < 8; ++s) {
        double q[4] = {rnd(), rnd(), rnd(), rnd()};
        double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        for (int i = 0; i < 4; ++i) q[i] /= n;
        const double w = q[0], x = q[1], y = q[2], z = q[3];
        double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                       2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                       2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        double *m = frames + 16 * s;
        memset(m, 0, 16 * sizeof(double));
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) m[4 * r + c] = R[3 * r + c]; m[4 * r + 3] = 0.3 * rnd(); }
        m[15] = 1.0;
    }
    double lim[7] = {0};
    chain_t c;
    chain_init(&c, frames, lim, lim);
    long mismatches = 0, flagged = 0, nonzero_status = 0;
    for (int t = 0; t < trials; ++t) {
        double J[6][7];
        if (t % 10 < 7) {
            double q[7];
            for (int a = 0; a < 7; ++a) q[a] = 3.0 * rnd();
            chain_jacobian(&c, q, J);
        } else {
            for (int r = 0; r < 6; ++r) for (int k = 0; k < 7; ++k) J[r][k] = rnd();
        }
        if (t % 10 == 7) for (int k = 0; k < 7; ++k) J[5][k] = J[4][k];                       // rank deficient
        if (t % 10 == 8) for (int k = 0; k < 7; ++k) { J[2][k] *= 1e-9; J[0][k] *= 1e6; }     // badly scaled
        if (t % 10 == 9) for (int r = 0; r < 6; ++r) { J[r][3] = 0.0; J[r][6] = J[r][1]; }    // zero / repeated columns
        double U0[6][7], w0[7], V0[7][7];
        const int rc0 = svd_hh(J, U0, w0, V0, 150);
        double U1[6][7], w1[7], tmp1[7], V1[49];
        memcpy(U1, J, sizeof(U1));
        omgb::VRef V{V1, 1};
        const int rc1 = omgb::r_svd(U1, w1, V, tmp1, 150);
        nonzero_status += rc0 != 0;
        bool same = rc0 == rc1 && memcmp(U0, U1, sizeof(U0)) == 0 && memcmp(w0, w1, sizeof(w0)) == 0;
        for (int r = 0; r < 7 && same; ++r)
            for (int k = 0; k < 7; ++k)
                if (memcmp(&V0[r][k], &V1[r * 7 + k], sizeof(double)) != 0) { same = false; break; }
        if (!same) {
            if (mismatches < 5) fprintf(stderr, "mismatch at trial %d (kind %d): rc %d vs %d\n", t, t % 10, rc0, rc1);
            ++mismatches;
        }
        (void)flagged;
    }
    printf("trials %d mismatches %ld nonzero_status %ld cancellation_branches %ld\n", trials, mismatches, nonzero_status,
           omgb_svd_trace_cancellations);
    return mismatches == 0 ? 0 : 1;
}
