// Batched inverse kinematics for goal-set construction: what the reference does with a 4-process pool
// of PyKDL solvers (omg/planner.py:16-87, 296-455 -> robot_pykdl.py:257-289 -> KDL ChainIkSolverPos_NR_JL +
// ChainIkSolverVel_pinv + SVD_HH, orocos_kdl/src/chainiksolverpos_nr_jl.cpp:61-101, chainiksolvervel_pinv.cpp:61-123,
// utilities/svd_HH.cpp:56-273) as ONE launch: one thread per (grasp pose, seed) runs that pair's whole chain of
// solves -- the standoff pose first, then every pose of the reach tail seeded with the previous solution
// (solve_one_pose_ik) -- entirely in registers / local memory.
//
// Same algorithm, same constants (100 Newton steps, 1e-6 twist tolerance, singular values < 1e-5 dropped, joint
// clamping after every step, KDL's Rot2 / GetRot / Norm formulas, the 0-or-1 `anorm` of svd_HH.cpp:136-138), fp64
// throughout.  This translation unit is compiled with -fmad=false so that sums of products round like the CPU code;
// what remains different from the reference is the last bit of sin / cos / acos.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/omgb200.h"
#include "host_common.h"
#include "ik_svd_reg.cuh"

namespace omgb {

constexpr int IK_NJ = 7, IK_NSEG = 8;

struct IkChain {               // kdl_parser's chain for panda_link0 -> panda_hand, built on the host
    double axis[IK_NSEG][3];   // joint axis in the parent frame, normalised (Joint's constructor)
    double origin[IK_NSEG][3];
    double tipM[IK_NSEG][9];   // f_tip relative to the joint: (R_parent_joint, 0)
    double qmin[IK_NJ], qmax[IK_NJ];
};

struct Fr { double M[9]; double p[3]; };

__device__ __forceinline__ void d_rot_mul(const double *a, const double *b, double *c) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
__device__ __forceinline__ void d_rot_vec(const double *a, const double *v, double *o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = a[3 * i] * v[0] + a[3 * i + 1] * v[1] + a[3 * i + 2] * v[2];
}
__device__ __forceinline__ void d_cross(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double d_sqr(double x) { return x * x; }
__device__ double d_vec_norm(const double *d) {   // frames.cpp:118-143
    double t1 = fabs(d[0]), t2 = fabs(d[1]);
    if (t1 >= t2) {
        t2 = fabs(d[2]);
        if (t1 >= t2) {
            if (t1 == 0) return 0;
            return t1 * sqrt(1 + d_sqr(d[1] / d[0]) + d_sqr(d[2] / d[0]));
        }
        return t2 * sqrt(1 + d_sqr(d[0] / d[2]) + d_sqr(d[1] / d[2]));
    }
    t1 = fabs(d[2]);
    if (t2 > t1) return t2 * sqrt(1 + d_sqr(d[0] / d[1]) + d_sqr(d[2] / d[1]));
    return t1 * sqrt(1 + d_sqr(d[0] / d[2]) + d_sqr(d[1] / d[2]));
}
__device__ __forceinline__ void d_rot2(const double *v, double angle, double *R) {   // frames.cpp:304-331
    double st, ct;
    sincos(angle, &st, &ct);
    const double vt = 1 - ct;
    const double m_vt_0 = vt * v[0], m_vt_1 = vt * v[1], m_vt_2 = vt * v[2];
    const double m_st_0 = v[0] * st, m_st_1 = v[1] * st, m_st_2 = v[2] * st;
    const double m_vt_0_1 = m_vt_0 * v[1], m_vt_0_2 = m_vt_0 * v[2], m_vt_1_2 = m_vt_1 * v[2];
    R[0] = ct + m_vt_0 * v[0];   R[1] = -m_st_2 + m_vt_0_1;   R[2] = m_st_1 + m_vt_0_2;
    R[3] = m_st_2 + m_vt_0_1;    R[4] = ct + m_vt_1 * v[1];   R[5] = -m_st_0 + m_vt_1_2;
    R[6] = -m_st_1 + m_vt_0_2;   R[7] = m_st_0 + m_vt_1_2;    R[8] = ct + m_vt_2 * v[2];
}

// Forward kinematics and Jacobian in one walk of the chain (chainfksolverpos_recursive.cpp and
// chainjnttojacsolver.cpp:49-95 compute the same frames with the same operations): tip frame `T`, J rows 0-2 linear /
// 3-5 angular, reference point at the tip.
template <bool UNROLL>
__device__ __forceinline__ void d_fk_jac(const IkChain &c, const double *q, Fr &T, double (*J)[IK_NJ], bool want_jac) {
#pragma unroll
    for (int i = 0; i < 9; ++i) T.M[i] = (i % 4 == 0) ? 1.0 : 0.0;
    T.p[0] = T.p[1] = T.p[2] = 0.0;
    if (want_jac) {
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int k = 0; k < IK_NJ; ++k) J[r][k] = 0.0;
    }
#pragma unroll(UNROLL ? IK_NSEG : 1)
    for (int s = 0; s < IK_NSEG; ++s) {
        const bool mov = s < IK_NJ;
        double PM[9], Pp[3];
        if (mov) {
            double jM[9];
            d_rot2(c.axis[s], q[s], jM);
            d_rot_mul(jM, c.tipM[s], PM);                       // joint.pose(q) * f_tip, f_tip = (R, 0)
            const double zero[3] = {0.0, 0.0, 0.0};
            double t[3];
            d_rot_vec(jM, zero, t);
#pragma unroll
            for (int i = 0; i < 3; ++i) Pp[i] = t[i] + c.origin[s][i];
        } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) PM[i] = c.tipM[s][i];   // Joint::None: identity * f_tip
#pragma unroll
            for (int i = 0; i < 3; ++i) Pp[i] = c.origin[s][i] + 0.0;
            // (identity.M * p + 0: each component is 1*p + 0*.. + 0*.. + 0, exact)
        }
        Fr total;
        d_rot_mul(T.M, PM, total.M);
        double t[3];
        d_rot_vec(T.M, Pp, t);
#pragma unroll
        for (int i = 0; i < 3; ++i) total.p[i] = t[i] + T.p[i];
        if (want_jac) {
            double d[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) d[i] = total.p[i] - T.p[i];
#pragma unroll
            for (int k = 0; k < IK_NJ; ++k) {                   // changeRefPoint of every column: vel += rot x d
                const double r[3] = {J[3][k], J[4][k], J[5][k]};
                double cr[3];
                d_cross(r, d, cr);
                J[0][k] = J[0][k] + cr[0]; J[1][k] = J[1][k] + cr[1]; J[2][k] = J[2][k] + cr[2];
            }
            if (mov) {
                // Segment::twist(q, 1) = (0 + axis x 0, axis), rotated into the base frame
                const double zero[3] = {0.0, 0.0, 0.0};
                double cr[3], jv[3], tv[3], tr[3];
                d_cross(c.axis[s], zero, cr);
                jv[0] = 0.0 + cr[0]; jv[1] = 0.0 + cr[1]; jv[2] = 0.0 + cr[2];
                d_rot_vec(T.M, jv, tv);
                d_rot_vec(T.M, c.axis[s], tr);
                J[0][s] = tv[0]; J[1][s] = tv[1]; J[2][s] = tv[2];
                J[3][s] = tr[0]; J[4][s] = tr[1]; J[5][s] = tr[2];
            }
        }
        T = total;
    }
}

__device__ __forceinline__ double d_pythag(double a, double b) {   // svd_HH.cpp:29-44
    const double at = fabs(a), bt = fabs(b);
    if (at > bt) { const double ct = bt / at; return at * sqrt(1.0 + ct * ct); }
    if (bt == 0) return 0.0;
    const double ct = at / bt;
    return bt * sqrt(1.0 + ct * ct);
}
__device__ __forceinline__ double d_sign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// svd_HH.cpp:56-273 on the 6 x 7 Jacobian, in place in U; returns 0 or -2 like the original.
__device__ int d_svd(double (*U)[IK_NJ], double *w, double (*V)[IK_NJ], double *tmp, int maxiter) {
    const int rows = 6, cols = IK_NJ;
    int i, its = -1, j, jj, k, nm = 0, ppi = 0;
    bool flag;
    double anorm = 0, c = 0, f = 0, h = 0, s = 0, scale = 0, x = 0, y = 0, z = 0, g = 0;
    for (i = 0; i < cols; i++) {                       // Householder reduction to bidiagonal form
        ppi = i + 1;
        tmp[i] = scale * g;
        g = s = scale = 0.0;
        if (i < rows) {
            for (k = i; k < rows; k++) scale += fabs(U[k][i]);
            if (scale != 0.0) {
                for (k = i; k < rows; k++) { U[k][i] /= scale; s += U[k][i] * U[k][i]; }
                f = U[i][i];
                g = -d_sign(sqrt(s), f);
                h = f * g - s;
                U[i][i] = f - g;
                for (j = ppi; j < cols; j++) {
                    for (s = 0.0, k = i; k < rows; k++) s += U[k][i] * U[k][j];
                    f = s / h;
                    for (k = i; k < rows; k++) U[k][j] += f * U[k][i];
                }
                for (k = i; k < rows; k++) U[k][i] *= scale;
            }
        }
        w[i] = scale * g;
        g = s = scale = 0.0;
        if ((i < rows) && (i + 1 != cols)) {
            for (k = ppi; k < cols; k++) scale += fabs(U[i][k]);
            if (scale != 0.0) {
                for (k = ppi; k < cols; k++) { U[i][k] /= scale; s += U[i][k] * U[i][k]; }
                f = U[i][ppi];
                g = -d_sign(sqrt(s), f);
                h = f * g - s;
                U[i][ppi] = f - g;
                for (k = ppi; k < cols; k++) tmp[k] = U[i][k] / h;
                for (j = ppi; j < rows; j++) {
                    for (s = 0.0, k = ppi; k < cols; k++) s += U[j][k] * U[i][k];
                    for (k = ppi; k < cols; k++) U[j][k] += s * tmp[k];
                }
                for (k = ppi; k < cols; k++) U[i][k] *= scale;
            }
        }
        {   // (sic) both operands of the max are bool in the original: anorm ends up 0 or 1
            const bool m1 = anorm != 0.0, m2 = (fabs(w[i]) + fabs(tmp[i])) != 0.0;
            anorm = (m1 || m2) ? 1.0 : 0.0;
        }
    }
    for (i = cols - 1; i >= 0; i--) {                  // accumulation of right-hand transformations
        if (i < cols - 1) {
            if (g != 0.0) {
                for (j = ppi; j < cols; j++) V[j][i] = (U[i][j] / U[i][ppi]) / g;
                for (j = ppi; j < cols; j++) {
                    for (s = 0.0, k = ppi; k < cols; k++) s += U[i][k] * V[k][j];
                    for (k = ppi; k < cols; k++) V[k][j] += s * V[k][i];
                }
            }
            for (j = ppi; j < cols; j++) V[i][j] = V[j][i] = 0.0;
        }
        V[i][i] = 1.0;
        g = tmp[i];
        ppi = i;
    }
    for (i = rows - 1; i >= 0; i--) {                  // accumulation of left-hand transformations (rows < cols)
        ppi = i + 1;
        g = w[i];
        for (j = ppi; j < cols; j++) U[i][j] = 0.0;
        if (g != 0.0) {
            g = 1.0 / g;
            for (j = ppi; j < cols; j++) {
                for (s = 0.0, k = ppi; k < rows; k++) s += U[k][i] * U[k][j];
                f = (s / U[i][i]) * g;
                for (k = i; k < rows; k++) U[k][j] += f * U[k][i];
            }
            for (j = i; j < rows; j++) U[j][i] *= g;
        } else {
            for (j = i; j < rows; j++) U[j][i] = 0.0;
        }
        U[i][i] = U[i][i] + 1.0;
    }
    for (k = cols - 1; k >= 0; k--) {                  // diagonalisation of the bidiagonal form
        for (its = 1; its <= maxiter; its++) {
            flag = true;
            for (ppi = k; ppi >= 0; ppi--) {
                nm = ppi - 1;
                if ((fabs(tmp[ppi]) + anorm) == anorm) { flag = false; break; }
                if (fabs(w[nm] + anorm) == anorm) break;
            }
            if (flag) {
                c = 0.0;
                s = 1.0;
                for (i = ppi; i <= k; i++) {
                    f = s * tmp[i];
                    tmp[i] = c * tmp[i];
                    if ((fabs(f) + anorm) == anorm) break;
                    g = w[i];
                    h = d_pythag(f, g);
                    w[i] = h;
                    h = 1.0 / h;
                    c = g * h;
                    s = (-f * h);
                    for (j = 0; j < rows; j++) {
                        y = U[j][nm]; z = U[j][i];
                        U[j][nm] = y * c + z * s;
                        U[j][i] = z * c - y * s;
                    }
                }
            }
            z = w[k];
            if (ppi == k) {
                if (z < 0.0) {
                    w[k] = -z;
                    for (j = 0; j < cols; j++) V[j][k] = -V[j][k];
                }
                break;
            }
            x = w[ppi];
            nm = k - 1;
            y = w[nm];
            g = tmp[nm];
            h = tmp[k];
            f = ((y - z) * (y + z) + (g - h) * (g + h)) / (2.0 * h * y);
            g = d_pythag(f, 1.0);
            f = ((x - z) * (x + z) + h * ((y / (f + d_sign(g, f))) - h)) / x;
            c = s = 1.0;
            for (j = ppi; j <= nm; j++) {
                i = j + 1;
                g = tmp[i];
                y = w[i];
                h = s * g;
                g = c * g;
                z = d_pythag(f, h);
                tmp[j] = z;
                c = f / z;
                s = h / z;
                f = x * c + g * s;
                g = g * c - x * s;
                h = y * s;
                y = y * c;
                for (jj = 0; jj < cols; jj++) {
                    x = V[jj][j]; z = V[jj][i];
                    V[jj][j] = x * c + z * s;
                    V[jj][i] = z * c - x * s;
                }
                z = d_pythag(f, h);
                w[j] = z;
                if (z != 0.0) { z = 1.0 / z; c = f * z; s = h * z; }
                f = (c * g) + (s * y);
                x = (c * y) - (s * g);
                for (jj = 0; jj < rows; jj++) {
                    y = U[jj][j]; z = U[jj][i];
                    U[jj][j] = y * c + z * s;
                    U[jj][i] = z * c - y * s;
                }
            }
            tmp[ppi] = 0.0;
            tmp[k] = f;
            w[k] = x;
        }
    }
    return its == maxiter ? -2 : 0;
}

// frames.cpp:337-431 (eps = KDL::epsilon = 1e-6): rotation vector axis * angle of a rotation matrix
__device__ void d_get_rot(const double *d, double *out) {
    const double eps = 0.000001, eps2 = eps * 10;
    double x, y, z;
    if ((fabs(d[1] - d[3]) < eps) && (fabs(d[2] - d[6]) < eps) && (fabs(d[5] - d[7]) < eps)) {
        if ((fabs(d[1] + d[3]) < eps2) && (fabs(d[2] + d[6]) < eps2) && (fabs(d[5] + d[7]) < eps2) &&
            (fabs(d[0] + d[4] + d[8] - 3) < eps2)) {
            out[0] = 0.0; out[1] = 0.0; out[2] = 0.0;
            return;
        }
        const double angle = 3.14159265358979323846;
        const double xx = (d[0] + 1) / 2, yy = (d[4] + 1) / 2, zz = (d[8] + 1) / 2;
        const double xy = (d[1] + d[3]) / 4, xz = (d[2] + d[6]) / 4, yz = (d[5] + d[7]) / 4;
        if ((xx > yy) && (xx > zz)) { x = sqrt(xx); y = xy / x; z = xz / x; }
        else if (yy > zz) { y = sqrt(yy); x = xy / y; z = yz / y; }
        else { z = sqrt(zz); x = xz / z; y = yz / z; }
        out[0] = x * angle; out[1] = y * angle; out[2] = z * angle;
        return;
    }
    const double f = (d[0] + d[4] + d[8] - 1) / 2;
    const double angle = acos(fmax(-1.0, fmin(1.0, f)));
    double a[3] = {d[7] - d[5], d[2] - d[6], d[3] - d[1]};
    const double v = d_vec_norm(a);
    if (v < eps) { a[0] = 1; a[1] = 0; a[2] = 0; }
    else { a[0] = a[0] / v; a[1] = a[1] / v; a[2] = a[2] / v; }
    out[0] = a[0] * angle; out[1] = a[1] * angle; out[2] = a[2] * angle;
}

// One ChainIkSolverPos_NR_JL::CartToJnt: q in/out.  Returns 0 (found), -5 (iteration limit) or -100 (SVD failure).
// REG: the factorisation keeps U / w / tmp in registers and V in shared memory (ik_svd_reg.cuh); otherwise the generic
// d_svd with everything in local memory.  Same operations in the same order either way.
template <bool REG>
__device__ int d_ik_solve(const IkChain &c, const double *target, double *q, int *steps, double *v_smem, int v_stride) {
    Fr goal;
    {
        const double x = target[3], y = target[4], z = target[5], w = target[6];
        const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
        goal.M[0] = w2 + x2 - y2 - z2;       goal.M[1] = 2 * x * y - 2 * w * z;   goal.M[2] = 2 * x * z + 2 * w * y;
        goal.M[3] = 2 * x * y + 2 * w * z;   goal.M[4] = w2 - x2 + y2 - z2;       goal.M[5] = 2 * y * z - 2 * w * x;
        goal.M[6] = 2 * x * z - 2 * w * y;   goal.M[7] = 2 * y * z + 2 * w * x;   goal.M[8] = w2 - x2 - y2 + z2;
        goal.p[0] = target[0]; goal.p[1] = target[1]; goal.p[2] = target[2];
    }
    const int maxiter = 100;
    const double eps = 1e-6, svd_eps = 0.00001;
    int i, status = 0;
    double U[6][IK_NJ], S[IK_NJ], tmp[IK_NJ];
    double Vloc[REG ? 1 : IK_NJ][IK_NJ];
    const VRef Vs{v_smem, v_stride};
    for (i = 0; i < maxiter; i++) {
        Fr f;
        d_fk_jac<REG>(c, q, f, U, true);
        double tw[6], Mi[9], Rrel[9], rv[3], rr[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) tw[a] = goal.p[a] - f.p[a];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) Mi[3 * a + b] = f.M[3 * b + a];
        d_rot_mul(Mi, goal.M, Rrel);
        d_get_rot(Rrel, rv);
        d_rot_vec(f.M, rv, rr);
        tw[3] = rr[0]; tw[4] = rr[1]; tw[5] = rr[2];
        bool zero = true;
#pragma unroll
        for (int a = 0; a < 6; ++a) zero = zero && (eps > tw[a]) && (tw[a] > -eps);
        if (zero) break;
        const int rc = REG ? r_svd(U, S, Vs, tmp, 150) : d_svd(U, S, Vloc, tmp, 150);
        if (rc != 0) { status = -100; break; }
#pragma unroll
        for (int a = 0; a < IK_NJ; ++a) {
            double sum = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) sum += U[b][a] * tw[b];
            tmp[a] = fabs(S[a]) < svd_eps ? 0.0 : sum / S[a];
        }
#pragma unroll
        for (int a = 0; a < IK_NJ; ++a) {
            double sum = 0.0;
#pragma unroll
            for (int b = 0; b < IK_NJ; ++b) sum += (REG ? Vs(a, b) : Vloc[REG ? 0 : a][b]) * tmp[b];
            double v = q[a] + sum;
            if (v < c.qmin[a]) v = c.qmin[a];
            if (v > c.qmax[a]) v = c.qmax[a];
            S[a] = v;   // (tmp is still being read; S is free)
        }
#pragma unroll
        for (int a = 0; a < IK_NJ; ++a) q[a] = S[a];
    }
    *steps = i;
    if (status) return status;
    return i != maxiter ? 0 : -5;
}

// targets [P,T,7] (position xyz, quaternion xyzw), seeds [S,7] -> sols [P,S,T,7], solved [P,S] (solves that succeeded
// before the first failure), steps [P,S,T] or null (Newton steps of every solve attempted).
template <bool REG>
__global__ void __launch_bounds__(64) ik_chain_kernel(const __grid_constant__ IkChain c,
                                                      const double *__restrict__ targets,
                                                      const double *__restrict__ seeds, int P, int T, int S,
                                                      double *__restrict__ sols, int *__restrict__ solved,
                                                      int *__restrict__ steps) {
    __shared__ double s_v[REG ? 49 * 64 : 1];   // V of every thread, [element][thread]
    const long long total = (long long)P * S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(idx % S);
        const long long p = idx / S;
        double q[IK_NJ];
#pragma unroll
        for (int a = 0; a < IK_NJ; ++a) q[a] = seeds[(size_t)s * IK_NJ + a];
        int t = 0;
        for (; t < T; ++t) {
            double tg[7];
#pragma unroll
            for (int a = 0; a < 7; ++a) tg[a] = targets[((size_t)p * T + t) * 7 + a];
            double r[IK_NJ];
#pragma unroll
            for (int a = 0; a < IK_NJ; ++a) r[a] = q[a];
            int st = 0;
            const int rc = d_ik_solve<REG>(c, tg, r, &st, s_v + (REG ? threadIdx.x : 0), 64);
            if (steps) steps[(size_t)idx * T + t] = st;
#pragma unroll
            for (int a = 0; a < IK_NJ; ++a) sols[((size_t)idx * T + t) * IK_NJ + a] = r[a];
            if (rc < 0) break;
#pragma unroll
            for (int a = 0; a < IK_NJ; ++a) q[a] = r[a];
        }
        solved[idx] = t;
    }
}

// hand pose by the same chain: joints [M,7] -> poses [M,16] row-major 4x4 (used by the goal-set filters)
__global__ void __launch_bounds__(128) ik_fk_kernel(const __grid_constant__ IkChain c, const double *__restrict__ joints,
                                                    long long joint_stride, int M, double *__restrict__ poses) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double q[IK_NJ];
#pragma unroll
    for (int a = 0; a < IK_NJ; ++a) q[a] = joints[(size_t)m * joint_stride + a];
    Fr f;
    d_fk_jac<false>(c, q, f, nullptr, false);
    double *o = poses + (size_t)m * 16;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < 3; ++k) o[4 * r + k] = f.M[3 * r + k];
        o[4 * r + 3] = f.p[r];
    }
    o[12] = o[13] = o[14] = 0.0; o[15] = 1.0;
}

static double h_vec_norm(const double *d) {
    double t1 = fabs(d[0]), t2 = fabs(d[1]);
    auto sq = [](double x) { return x * x; };
    if (t1 >= t2) {
        t2 = fabs(d[2]);
        if (t1 >= t2) {
            if (t1 == 0) return 0.0;
            return t1 * sqrt(1 + sq(d[1] / d[0]) + sq(d[2] / d[0]));
        }
        return t2 * sqrt(1 + sq(d[0] / d[2]) + sq(d[1] / d[2]));
    }
    t1 = fabs(d[2]);
    if (t2 > t1) return t2 * sqrt(1 + sq(d[0] / d[1]) + sq(d[2] / d[1]));
    return t1 * sqrt(1 + sq(d[0] / d[2]) + sq(d[1] / d[2]));
}

static int build_chain(const double *frames, const double *qmin, const double *qmax, IkChain *c, const char *who) {
    if (!frames) return host_fail(OMGB_ERR_INVALID, std::string(who) + ": null chain frames");
    memset(c, 0, sizeof(*c));
    for (int s = 0; s < IK_NSEG; ++s) {
        const double *m = frames + 16 * s;
        const double R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
        double a[3] = {R[0] * 0.0 + R[1] * 0.0 + R[2] * 1.0, R[3] * 0.0 + R[4] * 0.0 + R[5] * 1.0,
                       R[6] * 0.0 + R[7] * 0.0 + R[8] * 1.0};
        const double nrm = h_vec_norm(a);
        if (!(nrm > 0.0)) return host_fail(OMGB_ERR_INVALID, std::string(who) + ": degenerate joint frame");
        for (int i = 0; i < 3; ++i) { c->axis[s][i] = a[i] / nrm; c->origin[s][i] = m[4 * i + 3]; }
        memcpy(c->tipM[s], R, sizeof(R));
    }
    for (int j = 0; j < IK_NJ; ++j) { c->qmin[j] = qmin ? qmin[j] : -1e300; c->qmax[j] = qmax ? qmax[j] : 1e300; }
    return OMGB_OK;
}

}  // namespace omgb

using namespace omgb;

extern "C" int omgb_ik_solve(const double *chain_frames, const double *q_min, const double *q_max,
                             const double *d_targets, int num_poses, int chain_length, const double *d_seeds,
                             int num_seeds, double *d_sols, int *d_solved, int *d_steps, void *stream) {
    if (num_poses < 0 || num_seeds < 0 || chain_length < 1)
        return host_fail(OMGB_ERR_INVALID, "omgb_ik_solve: sizes must be non-negative, chain_length >= 1");
    if (!q_min || !q_max) return host_fail(OMGB_ERR_INVALID, "omgb_ik_solve: joint limits required");
    IkChain c;
    int rc = build_chain(chain_frames, q_min, q_max, &c, "omgb_ik_solve");
    if (rc) return rc;
    const long long total = (long long)num_poses * num_seeds;
    if (total == 0) return OMGB_OK;
    if (!d_targets || !d_seeds || !d_sols || !d_solved) return host_fail(OMGB_ERR_INVALID, "omgb_ik_solve: null buffer");
    long long blocks = (total + 63) / 64;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    // Two builds of the same arithmetic (identical results, tests/test_cpu_ik_svd_reg.py): the register-resident
    // factorisation has the shorter dependent chain but needs 255 registers and 25 KB of shared memory per 64 chains,
    // the generic one keeps U / V in local memory at 168 registers.  Measured on B200: 3900 chains 40.8 vs 44.5 ms,
    // 39000 chains 189 vs 78 ms -- so goal-set sized problems take the first, up-sampled sets the second.
    // OMGB_IK_SVD=reg|local forces one.
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("OMGB_IK_SVD");
        forced = !e ? -1 : (strcmp(e, "local") == 0 ? 0 : (strcmp(e, "reg") == 0 ? 1 : -1));
    }
    const int use_reg = forced >= 0 ? forced : (total <= 4096 ? 1 : 0);
    if (use_reg)
        ik_chain_kernel<true><<<(int)blocks, 64, 0, (cudaStream_t)stream>>>(c, d_targets, d_seeds, num_poses,
                                                                           chain_length, num_seeds, d_sols, d_solved,
                                                                           d_steps);
    else
        ik_chain_kernel<false><<<(int)blocks, 64, 0, (cudaStream_t)stream>>>(c, d_targets, d_seeds, num_poses,
                                                                            chain_length, num_seeds, d_sols, d_solved,
                                                                            d_steps);
    host_count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return host_fail(OMGB_ERR_CUDA, std::string("ik_chain_kernel: ") + cudaGetErrorString(e));
    return OMGB_OK;
}

extern "C" int omgb_hand_poses(const double *chain_frames, const double *d_joints, long long joint_stride,
                               int num_configs, double *d_poses, void *stream) {
    if (num_configs < 0 || joint_stride < IK_NJ)
        return host_fail(OMGB_ERR_INVALID, "omgb_hand_poses: num_configs >= 0 and joint_stride >= 7 required");
    IkChain c;
    int rc = build_chain(chain_frames, nullptr, nullptr, &c, "omgb_hand_poses");
    if (rc) return rc;
    if (num_configs == 0) return OMGB_OK;
    if (!d_joints || !d_poses) return host_fail(OMGB_ERR_INVALID, "omgb_hand_poses: null buffer");
    ik_fk_kernel<<<(num_configs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c, d_joints, joint_stride, num_configs,
                                                                            d_poses);
    host_count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return host_fail(OMGB_ERR_CUDA, std::string("ik_fk_kernel: ") + cudaGetErrorString(e));
    return OMGB_OK;
}
