// SVD_HH (orocos_kdl/src/utilities/svd_HH.cpp:56-273) for the 6 x 7 Jacobian with every index of U, w and tmp known at
// compile time, so that the 56 doubles of U / w / tmp live in registers (the generic version keeps them in local
// memory and spends 29 % of its instructions on loads and stores inside dependent chains).  V sits in shared memory
// ([element][thread], conflict free).  The operations and their order are exactly those of the generic version
// (d_svd in ik_kernels.cu), which is what tests/host_svd_check.cpp verifies bit for bit on the host against
// oracle/kdl_ik_ref.c; loops whose bounds depend on data run over their full static range under a predicate.
// The rare "cancellation" branch (a negligible singular value above the split) indexes U by a runtime column: it works
// on a scratch copy.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define OMGB_HD __host__ __device__ __forceinline__
#else
#define OMGB_HD inline
#endif

#ifdef OMGB_SVD_TRACE
static long omgb_svd_trace_cancellations = 0;   // host test instrumentation
#endif

namespace omgb {

OMGB_HD double r_pythag(double a, double b) {
    const double at = fabs(a), bt = fabs(b);
    if (at > bt) { const double ct = bt / at; return at * sqrt(1.0 + ct * ct); }
    if (bt == 0) return 0.0;
    const double ct = at / bt;
    return bt * sqrt(1.0 + ct * ct);
}
OMGB_HD double r_sign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

OMGB_HD double r_get7(const double (&a)[7], int i) {
    switch (i) {
        case 0: return a[0]; case 1: return a[1]; case 2: return a[2]; case 3: return a[3];
        case 4: return a[4]; case 5: return a[5]; default: return a[6];
    }
}
OMGB_HD void r_set7(double (&a)[7], int i, double v) {
    switch (i) {
        case 0: a[0] = v; break; case 1: a[1] = v; break; case 2: a[2] = v; break; case 3: a[3] = v; break;
        case 4: a[4] = v; break; case 5: a[5] = v; break; default: a[6] = v; break;
    }
}

// V accessor: element (r, c) of this thread's 7 x 7 matrix at base[(r * 7 + c) * stride]
struct VRef {
    double *base;
    int stride;
    OMGB_HD double &operator()(int r, int c) const { return base[(r * 7 + c) * stride]; }
};

// One singular value K of the diagonalisation (svd_HH.cpp:190-268).  Returns the `its` the loop ended with.
template <int K>
OMGB_HD int r_qr_value(double (&U)[6][7], double (&w)[7], double (&tmp)[7], const VRef &V, double anorm, int maxiter) {
    int its;
    double c, f, h, s, x, y, z, g;
    for (its = 1; its <= maxiter; its++) {
        bool flag = true;
        int ppi = K, nm = K - 1;
        {   // test for splitting: ppi runs down from K
            bool found = false;
#pragma unroll
            for (int p = K; p >= 0; p--) {
                if (!found) {
                    ppi = p; nm = p - 1;
                    if ((fabs(tmp[p]) + anorm) == anorm) { flag = false; found = true; }
                    else if (p > 0 && (fabs(w[p - 1] + anorm) == anorm)) { found = true; }
                    // (p == 0 always takes the first exit: tmp[0] is zero)
                }
            }
        }
        if (flag) {
#ifdef OMGB_SVD_TRACE
            ++omgb_svd_trace_cancellations;
#endif
            c = 0.0;
            s = 1.0;
            if (ppi == K) {
                // the usual case (a negligible singular value right above the last one: every 6 x 7 Jacobian has a
                // null direction): the loop below runs once, with i = K and nm = K - 1 -- static columns
                constexpr int KM = K > 0 ? K - 1 : 0;
                f = s * tmp[K];
                tmp[K] = c * tmp[K];
                if (!((fabs(f) + anorm) == anorm)) {
                    g = w[K];
                    h = r_pythag(f, g);
                    w[K] = h;
                    h = 1.0 / h;
                    c = g * h;
                    s = (-f * h);
#pragma unroll
                    for (int j = 0; j < 6; j++) {
                        y = U[j][KM]; z = U[j][K];
                        U[j][KM] = y * c + z * s;
                        U[j][K] = z * c - y * s;
                    }
                }
            } else {
                // split further up (not observed on arm Jacobians): runtime columns -> work on a scratch copy of U
                double Ut[6][7];
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int q = 0; q < 7; ++q) Ut[r][q] = U[r][q];
                for (int i = ppi; i <= K; i++) {
                    f = s * r_get7(tmp, i);
                    r_set7(tmp, i, c * r_get7(tmp, i));
                    if ((fabs(f) + anorm) == anorm) break;
                    g = r_get7(w, i);
                    h = r_pythag(f, g);
                    r_set7(w, i, h);
                    h = 1.0 / h;
                    c = g * h;
                    s = (-f * h);
                    for (int j = 0; j < 6; j++) {
                        y = Ut[j][nm]; z = Ut[j][i];
                        Ut[j][nm] = y * c + z * s;
                        Ut[j][i] = z * c - y * s;
                    }
                }
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int q = 0; q < 7; ++q) U[r][q] = Ut[r][q];
            }
        }
        z = w[K];
        if (ppi == K) {   // convergence
            if (z < 0.0) {
                w[K] = -z;
#pragma unroll
                for (int j = 0; j < 7; j++) V(j, K) = -V(j, K);
            }
            break;
        }
        if (K > 0) {
            x = r_get7(w, ppi);
            y = w[K > 0 ? K - 1 : 0];
            g = tmp[K > 0 ? K - 1 : 0];
            h = tmp[K];
            f = ((y - z) * (y + z) + (g - h) * (g + h)) / (2.0 * h * y);
            g = r_pythag(f, 1.0);
            f = ((x - z) * (x + z) + h * ((y / (f + r_sign(g, f))) - h)) / x;
            c = s = 1.0;
#pragma unroll
            for (int j = 0; j < (K > 0 ? K : 1); j++) {   // j = ppi .. K-1
                if (K > 0 && j >= ppi) {
                    constexpr int dummy = 0; (void)dummy;
                    const int i = j + 1;
                    g = tmp[i < 7 ? i : 6];
                    y = w[i < 7 ? i : 6];
                    h = s * g;
                    g = c * g;
                    z = r_pythag(f, h);
                    tmp[j] = z;
                    c = f / z;
                    s = h / z;
                    f = x * c + g * s;
                    g = g * c - x * s;
                    h = y * s;
                    y = y * c;
#pragma unroll
                    for (int jj = 0; jj < 7; jj++) {
                        x = V(jj, j); z = V(jj, i);
                        V(jj, j) = x * c + z * s;
                        V(jj, i) = z * c - x * s;
                    }
                    z = r_pythag(f, h);
                    w[j] = z;
                    if (z != 0.0) { z = 1.0 / z; c = f * z; s = h * z; }
                    f = (c * g) + (s * y);
                    x = (c * y) - (s * g);
#pragma unroll
                    for (int jj = 0; jj < 6; jj++) {
                        y = U[jj][j]; z = U[jj][i < 7 ? i : 6];
                        U[jj][j] = y * c + z * s;
                        U[jj][i < 7 ? i : 6] = z * c - y * s;
                    }
                }
            }
            r_set7(tmp, ppi, 0.0);
            tmp[K] = f;
            w[K] = x;
        }
    }
    return its;
}

// U holds the Jacobian on entry (6 x 7) and the left vectors on exit; returns 0 or -2 like the original.
OMGB_HD int r_svd(double (&U)[6][7], double (&w)[7], const VRef &V, double (&tmp)[7], int maxiter) {
    constexpr int rows = 6, cols = 7;
    double anorm = 0, f = 0, h = 0, s = 0, scale = 0, g = 0;
    // ---- Householder reduction to bidiagonal form ------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < cols; i++) {
        const int ppi = i + 1;
        tmp[i] = scale * g;
        g = s = scale = 0.0;
        if (i < rows) {
#pragma unroll
            for (int k = 0; k < rows; k++) if (k >= i) scale += fabs(U[k][i]);
            if (scale != 0.0) {
#pragma unroll
                for (int k = 0; k < rows; k++) if (k >= i) { U[k][i] /= scale; s += U[k][i] * U[k][i]; }
                f = U[i < rows ? i : 0][i];
                g = -r_sign(sqrt(s), f);
                h = f * g - s;
                U[i < rows ? i : 0][i] = f - g;
#pragma unroll
                for (int j = 0; j < cols; j++) if (j >= ppi) {
                    s = 0.0;
#pragma unroll
                    for (int k = 0; k < rows; k++) if (k >= i) s += U[k][i] * U[k][j];
                    f = s / h;
#pragma unroll
                    for (int k = 0; k < rows; k++) if (k >= i) U[k][j] += f * U[k][i];
                }
#pragma unroll
                for (int k = 0; k < rows; k++) if (k >= i) U[k][i] *= scale;
            }
        }
        w[i] = scale * g;
        g = s = scale = 0.0;
        if ((i < rows) && (i + 1 != cols)) {
            constexpr int dummy = 0; (void)dummy;
            const int ir = i < rows ? i : 0;
#pragma unroll
            for (int k = 0; k < cols; k++) if (k >= ppi) scale += fabs(U[ir][k]);
            if (scale != 0.0) {
#pragma unroll
                for (int k = 0; k < cols; k++) if (k >= ppi) { U[ir][k] /= scale; s += U[ir][k] * U[ir][k]; }
                f = U[ir][ppi < cols ? ppi : 0];
                g = -r_sign(sqrt(s), f);
                h = f * g - s;
                U[ir][ppi < cols ? ppi : 0] = f - g;
#pragma unroll
                for (int k = 0; k < cols; k++) if (k >= ppi) tmp[k] = U[ir][k] / h;
#pragma unroll
                for (int j = 0; j < rows; j++) if (j >= ppi) {
                    s = 0.0;
#pragma unroll
                    for (int k = 0; k < cols; k++) if (k >= ppi) s += U[j][k] * U[ir][k];
#pragma unroll
                    for (int k = 0; k < cols; k++) if (k >= ppi) U[j][k] += s * tmp[k];
                }
#pragma unroll
                for (int k = 0; k < cols; k++) if (k >= ppi) U[ir][k] *= scale;
            }
        }
        {   // (sic) both operands of the max are bool in the original: anorm ends up 0 or 1
            const bool m1 = anorm != 0.0, m2 = (fabs(w[i]) + fabs(tmp[i])) != 0.0;
            anorm = (m1 || m2) ? 1.0 : 0.0;
        }
    }
    // ---- accumulation of right-hand transformations (g = 0 after the last pass above, ppi = cols) ----------------------
#pragma unroll
    for (int i = cols - 1; i >= 0; i--) {
        const int ppi = i + 1;
        if (i < cols - 1) {
            const int ir = i < rows ? i : 0;   // (i <= 5 here)
            if (g != 0.0) {
#pragma unroll
                for (int j = 0; j < cols; j++) if (j >= ppi) V(j, i) = (U[ir][j] / U[ir][ppi < cols ? ppi : 0]) / g;
#pragma unroll
                for (int j = 0; j < cols; j++) if (j >= ppi) {
                    s = 0.0;
#pragma unroll
                    for (int k = 0; k < cols; k++) if (k >= ppi) s += U[ir][k] * V(k, j);
#pragma unroll
                    for (int k = 0; k < cols; k++) if (k >= ppi) V(k, j) += s * V(k, i);
                }
            }
#pragma unroll
            for (int j = 0; j < cols; j++) if (j >= ppi) { V(i, j) = 0.0; V(j, i) = 0.0; }
        }
        V(i, i) = 1.0;
        g = tmp[i];
    }
    // ---- accumulation of left-hand transformations -----------------------------------------------------------------
#pragma unroll
    for (int i = rows - 1; i >= 0; i--) {
        const int ppi = i + 1;
        g = w[i];
#pragma unroll
        for (int j = 0; j < cols; j++) if (j >= ppi) U[i][j] = 0.0;
        if (g != 0.0) {
            g = 1.0 / g;
#pragma unroll
            for (int j = 0; j < cols; j++) if (j >= ppi) {
                s = 0.0;
#pragma unroll
                for (int k = 0; k < rows; k++) if (k >= ppi) s += U[k][i] * U[k][j];
                f = (s / U[i][i]) * g;
#pragma unroll
                for (int k = 0; k < rows; k++) if (k >= i) U[k][j] += f * U[k][i];
            }
#pragma unroll
            for (int j = 0; j < rows; j++) if (j >= i) U[j][i] *= g;
        } else {
#pragma unroll
            for (int j = 0; j < rows; j++) if (j >= i) U[j][i] = 0.0;
        }
        U[i][i] = U[i][i] + 1.0;
    }
    // ---- diagonalisation of the bidiagonal form ------------------------------------------------------------------------
    int its;
    its = r_qr_value<6>(U, w, tmp, V, anorm, maxiter);
    its = r_qr_value<5>(U, w, tmp, V, anorm, maxiter);
    its = r_qr_value<4>(U, w, tmp, V, anorm, maxiter);
    its = r_qr_value<3>(U, w, tmp, V, anorm, maxiter);
    its = r_qr_value<2>(U, w, tmp, V, anorm, maxiter);
    its = r_qr_value<1>(U, w, tmp, V, anorm, maxiter);
    its = r_qr_value<0>(U, w, tmp, V, anorm, maxiter);
    return its == maxiter ? -2 : 0;
}

}  // namespace omgb
