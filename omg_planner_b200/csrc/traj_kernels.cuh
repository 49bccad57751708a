// Trajectory initialisation on the device: omg/util.py:238-258 interpolate_waypoints, the routine behind
// Trajectory.interpolate_waypoints (omg/core.py:59-78), for a whole batch in one launch.
//
//   knots    x_k = linspace(0, 1, K)           (waypoints [B,K,9])
//   samples  t_i = linspace(0, 1, n + 2)[1:-1] (the n interior points)
//   mode 1   scipy.interpolate.CubicSpline(x, y, bc_type="clamped"): first derivative 0 at both ends; the knot
//            derivatives s solve the tridiagonal system of scipy/interpolate/_cubic.py (rows 1..K-2:
//            dx_i s_{i-1} + 2 (dx_{i-1} + dx_i) s_i + dx_{i-1} s_{i+1} = 3 (dx_i slope_{i-1} + dx_{i-1} slope_i)),
//            the piece coefficients are CubicHermiteSpline's, the value is PPoly's power sum c3 + c2 s + c1 s^2 + c0 s^3.
//   mode 0   scipy.interpolate.interp1d(x, y, "linear"): the convex combination
//            (t - x_lo)/(x_hi - x_lo) * y_hi + (x_hi - t)/(x_hi - x_lo) * y_lo.
//
// One thread per output element (coalesced fp64 stores); K is small (2 in every call the reference makes), so each
// thread re-solves its own (trajectory, DOF) system in registers rather than staging it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace omgb {

constexpr int TRAJ_MAX_KNOTS = 32;

__device__ __forceinline__ double linspace_at(int k, int num) {   // numpy.linspace(0, 1, num)[k]
    if (k == num - 1) return 1.0;
    const double step = 1.0 / (double)(num - 1);
    return (double)k * step;
}

__global__ void __launch_bounds__(256) traj_interpolate_kernel(const double *__restrict__ waypoints, int batch, int K,
                                                               int n, int mode, double *__restrict__ xi) {
    const long long total = (long long)batch * n * 9;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(idx % 9);
        const long long bi = idx / 9;
        const int i = (int)(bi % n);
        const long long b = bi / n;
        const double *y = waypoints + (size_t)b * K * 9 + d;   // y[k] = y[k * 9]
        const double t = linspace_at(i + 1, n + 2);
        // interval: x[lo] <= t < x[lo + 1] (PPoly / searchsorted agree for interior samples)
        int lo = 0;
        for (int k = 1; k < K - 1; ++k)
            if (linspace_at(k, K) <= t) lo = k;
        const double x_lo = linspace_at(lo, K), x_hi = linspace_at(lo + 1, K);
        const double y_lo = y[(size_t)lo * 9], y_hi = y[(size_t)(lo + 1) * 9];
        if (mode == 0) {
            xi[idx] = __dadd_rn(__dmul_rn(__ddiv_rn(t - x_lo, x_hi - x_lo), y_hi),
                                __dmul_rn(__ddiv_rn(x_hi - t, x_hi - x_lo), y_lo));
            continue;
        }
        double s_lo = 0.0, s_hi = 0.0;
        if (K > 2) {
            // Thomas elimination of the (diagonally dominant) clamped system; rows 0 and K-1 are s = 0
            double cp[TRAJ_MAX_KNOTS], dp[TRAJ_MAX_KNOTS];
            cp[0] = 0.0; dp[0] = 0.0;
            for (int k = 1; k < K - 1; ++k) {
                const double dx0 = linspace_at(k, K) - linspace_at(k - 1, K);
                const double dx1 = linspace_at(k + 1, K) - linspace_at(k, K);
                const double sl0 = (y[(size_t)k * 9] - y[(size_t)(k - 1) * 9]) / dx0;
                const double sl1 = (y[(size_t)(k + 1) * 9] - y[(size_t)k * 9]) / dx1;
                const double lower = dx1, diag = 2.0 * (dx0 + dx1), upper = dx0;
                const double rhs = 3.0 * (dx1 * sl0 + dx0 * sl1);
                const double den = diag - lower * cp[k - 1];
                cp[k] = upper / den;
                dp[k] = (rhs - lower * dp[k - 1]) / den;
            }
            double s_next = 0.0;   // s[K-1]
            for (int k = K - 2; k >= 1; --k) {
                const double sk = dp[k] - cp[k] * s_next;
                if (k == lo + 1) s_hi = sk;
                if (k == lo) s_lo = sk;
                s_next = sk;
            }
        }
        const double dx = x_hi - x_lo;
        const double slope = __ddiv_rn(y_hi - y_lo, dx);
        const double tt = __ddiv_rn(__dsub_rn(__dadd_rn(s_lo, s_hi), __dmul_rn(2.0, slope)), dx);
        const double c0 = __ddiv_rn(tt, dx);
        const double c1 = __dsub_rn(__ddiv_rn(__dsub_rn(slope, s_lo), dx), tt);
        const double c2 = s_lo, c3 = y_lo;
        const double s = t - x_lo;
        double res = c3, z = s;
        res = __dadd_rn(res, __dmul_rn(c2, z)); z = __dmul_rn(z, s);
        res = __dadd_rn(res, __dmul_rn(c1, z)); z = __dmul_rn(z, s);
        res = __dadd_rn(res, __dmul_rn(c0, z));
        xi[idx] = res;
    }
}

}  // namespace omgb
