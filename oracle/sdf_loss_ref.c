/*
 * ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * CPU restatement (plain C, fp32, single thread) of the reference CUDA operator
 *   omg_cuda.sdf_loss_forward  ==  layers/sdf_matching_loss_kernel.cu:204-262
 * following, line by line:
 *   lerp                      kernel.cu:15-18    a + t*(b-a)
 *   getValue index order      kernel.cu:30-34    x*Y*Z + y*Z + z
 *   getValueInterpolated      kernel.cu:37-64    (p-0.5) in double, (int) truncation, any OOB tap -> 1.0
 *   getGradientInterpolated   kernel.cu:67-86    six re-interpolations at p +/- e_k, 0.5*(f+ - f-)/delta
 *   SDFdistanceForward        kernel.cu:97-181   disabled -> transform -> value -> collide -> gradient ->
 *                                                potential branches -> rotate back
 *   sum_gradients             kernel.cu:186-195  sum over objects (reference: atomicAdd, order
 *                                                nondeterministic; here ascending object order)
 *
 * Third-party arithmetic that is NOT under /root/reference: Eigen (un-vendored, unpinned HEAD of
 * eigenteam/eigen-git-mirror per docker/install_deps.sh:56-64).  The kernel builds
 * Sophus::SE3<float>(Matrix4) (Sophus/sophus/se3.hpp:387-389 -> so3.hpp:392: Eigen::Quaternion(Matrix3)),
 * rotates points with Eigen's Quaternion::_transformVector (so3.hpp:298-300) and rotates the gradient
 * back with Quaternion::toRotationMatrix()^T (kernel.cu:126,176).  Those three Eigen routines are
 * restated here from Eigen 3.3's published Quaternion.h algorithms (Shoemake matrix->quaternion;
 * v + w*(2 q x v) + q x (2 q x v); the standard 12-product quaternion->matrix).
 *
 * Floating-point contraction: the reference is compiled by nvcc, which fuses a*b+c into FMA where its
 * optimiser likes.  The pattern restated here (fmaf() marks every fused op; everything else is a separately
 * rounded IEEE op; compile with -ffp-contract=off) is the one nvcc 12.9 -- the toolchain of this image, and the
 * oldest that targets sm_100 -- emits when it compiles the reference's own source for sm_100a with the flags of
 * layers/setup.py: oracle/sdf_ref/ builds that source where it lies (over stand-ins for the absent
 * ATen/Eigen/Sophus headers) and its SASS was read instruction by instruction.  The double-precision steps of the
 * reference ((p-0.5) in double; 0.5*(f+-f-)/delta in double; -v+0.5*eps in double) are evaluated in
 * double here exactly as written.
 *
 * PARITY STATUS: the reference ships no golden vectors for this operator.  Pinned by
 * (a) the reference's own getValueInterpolated / getGradientInterpolated (kernel.cu:37-86) compiled from the
 *     reference source for the host (oracle/_ref/libsdf_ref.so): bit-exact on random grid coordinates incl. the
 *     (-0.5, 0.5) truncation band, border voxels and out-of-bounds taps (tests/test_oracle_sdf.py, fixture
 *     tests/golden/sdf_interp.npz);
 * (b) the reference's own device path (sdf_loss_cuda_forward with its kernels, same library) run on the B200
 *     against the product operator, which is bit-identical to this file (tests/test_gpu_ref_operator.py);
 * (c) analytic SDF fields whose trilinear interpolant is known in closed form;
 * (d) agreement of the full CHOMP step built on it with the reference's own Python (tests/golden).
 * Not pinnable here: Eigen itself (absent, unpinned upstream) -- its three quaternion routines and the order of
 * its 3-term reductions are restated from Eigen 3.3's published code in oracle/sdf_ref/shim/Eigen/Core.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

typedef struct { float x, y, z; } f3;
typedef struct { int x, y, z; } i3;

static inline float lerpf(float a, float b, float t) { return fmaf(t, b - a, a); } /* kernel.cu:15-18 */

static inline float tap(const float *g, i3 dim, int x, int y, int z) {       /* kernel.cu:30-34 */
    return g[(size_t)x * dim.y * dim.z + (size_t)y * dim.z + z];
}

/* kernel.cu:37-64 */
static float value_interp(f3 p, i3 dim, const float *g, int *inb) {
    const int x0 = (int)((double)p.x - 0.5); const float fx = (float)(((double)p.x - 0.5) - x0);
    const int y0 = (int)((double)p.y - 0.5); const float fy = (float)(((double)p.y - 0.5) - y0);
    const int z0 = (int)((double)p.z - 0.5); const float fz = (float)(((double)p.z - 0.5) - z0);
    const int x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
    if (!(x0 >= 0 && x1 < dim.x && y0 >= 0 && y1 < dim.y && z0 >= 0 && z1 < dim.z)) {
        if (inb) *inb = 0;
        return 1.0f;
    }
    if (inb) *inb = 1;
    const float dx00 = lerpf(tap(g, dim, x0, y0, z0), tap(g, dim, x1, y0, z0), fx);
    const float dx01 = lerpf(tap(g, dim, x0, y0, z1), tap(g, dim, x1, y0, z1), fx);
    const float dx10 = lerpf(tap(g, dim, x0, y1, z0), tap(g, dim, x1, y1, z0), fx);
    const float dx11 = lerpf(tap(g, dim, x0, y1, z1), tap(g, dim, x1, y1, z1), fx);
    const float dxy0 = lerpf(dx00, dx10, fy);
    const float dxy1 = lerpf(dx01, dx11, fy);
    return lerpf(dxy0, dxy1, fz);
}

/* kernel.cu:67-86 */
static f3 grad_interp(f3 p, i3 dim, const float *g, float delta) {
    f3 q, r;
    q = p; q.x = p.x + 1.0f; const float f_px = value_interp(q, dim, g, 0);
    q = p; q.y = p.y + 1.0f; const float f_py = value_interp(q, dim, g, 0);
    q = p; q.z = p.z + 1.0f; const float f_pz = value_interp(q, dim, g, 0);
    q = p; q.x = p.x - 1.0f; const float f_mx = value_interp(q, dim, g, 0);
    q = p; q.y = p.y - 1.0f; const float f_my = value_interp(q, dim, g, 0);
    q = p; q.z = p.z - 1.0f; const float f_mz = value_interp(q, dim, g, 0);
    r.x = (float)(0.5 * (double)(f_px - f_mx) / (double)delta);
    r.y = (float)(0.5 * (double)(f_py - f_my) / (double)delta);
    r.z = (float)(0.5 * (double)(f_pz - f_mz) / (double)delta);
    return r;
}

/* Eigen::Quaternion<float>(Matrix3f) -- QuaternionBase::operator=(MatrixBase), Eigen 3.3 Quaternion.h */
static void mat_to_quat(const float m[3][3], float *qw, float qv[3]) {
    float t = m[0][0] + (m[1][1] + m[2][2]);   /* trace() = diagonal().sum(): unrolled tree redux c0 + (c1 + c2) */
    if (t > 0.0f) {
        t = sqrtf(t + 1.0f);
        *qw = 0.5f * t;
        t = 0.5f / t;
        qv[0] = (m[2][1] - m[1][2]) * t;
        qv[1] = (m[0][2] - m[2][0]) * t;
        qv[2] = (m[1][0] - m[0][1]) * t;
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrtf(m[i][i] - m[j][j] - m[k][k] + 1.0f);
        qv[i] = 0.5f * t;
        t = 0.5f / t;
        *qw = (m[k][j] - m[j][k]) * t;
        qv[j] = (m[j][i] + m[i][j]) * t;
        qv[k] = (m[k][i] + m[i][k]) * t;
    }
}

/* Eigen::QuaternionBase::toRotationMatrix (tx = 2x ...; twx = tx*w ...; res(0,1) = txy - twz ...), with the
 * multiply-adds fused exactly where nvcc 12.9 fuses them when it compiles the reference source for sm_100a
 * (oracle/sdf_ref: SASS of SDFdistanceForward; verified on the B200 by tests/test_gpu_ref_operator.py):
 * rounded products twx, twz, txz, tyy, tzz; txx, txy, twy, tyz exist only inside an FMA. */
static void quat_to_mat(float w, const float v[3], float R[3][3]) {
    const float tx = 2.0f * v[0], ty = 2.0f * v[1], tz = 2.0f * v[2];
    const float twx = tx * w, twz = tz * w, txz = tz * v[0], tyy = ty * v[1], tzz = tz * v[2];
    R[0][0] = 1.0f - (tyy + tzz);            R[0][1] = fmaf(ty, v[0], -twz);          R[0][2] = fmaf(ty, w, txz);
    R[1][0] = fmaf(ty, v[0], twz);           R[1][1] = 1.0f - fmaf(tx, v[0], tzz);    R[1][2] = fmaf(tz, v[1], -twx);
    R[2][0] = fmaf(-ty, w, txz);             R[2][1] = fmaf(tz, v[1], twx);           R[2][2] = 1.0f - fmaf(tx, v[0], tyy);
}

static inline f3 cross_fma(const float a[3], f3 b) {
    f3 r;
    r.x = fmaf(a[1], b.z, -(a[2] * b.y));
    r.y = fmaf(a[2], b.x, -(a[0] * b.z));
    r.z = fmaf(a[0], b.y, -(a[1] * b.x));
    return r;
}

/*
 * C entry point (mirrors the tensor contract of layers/omg_layers.cpp:24-49).
 *   pose_init [O,4,4] sdf_grids [O,X,Y,Z] sdf_limits [O,10] points [N,3]
 *   epsilons/padding_scales/clearances/disables [O]
 *   out: potentials [N], potential_grads [N,3], collides [N]; returns the number of in-bounds
 *   (point, enabled object) pairs P_in (SURVEY.md 8d algorithmic-bytes model).
 */
long long omg_oracle_sdf_loss(const float *pose_init, const float *sdf_grids, const float *sdf_limits,
                              const float *points, const float *epsilons, const float *padding_scales,
                              const float *clearances, const float *disables, int num_points,
                              int num_objects, float *potentials, float *potential_grads, float *collides) {
    long long p_in = 0;
    for (int n = 0; n < num_points; ++n) {
        potentials[n] = 0.0f; collides[n] = 0.0f;
        potential_grads[3 * n] = potential_grads[3 * n + 1] = potential_grads[3 * n + 2] = 0.0f;
    }
    for (int o = 0; o < num_objects; ++o) {
        if (disables[o] > 0) continue;                                   /* kernel.cu:115 */
        const float *P = pose_init + 16 * o;
        float m[3][3] = {{P[0], P[1], P[2]}, {P[4], P[5], P[6]}, {P[8], P[9], P[10]}};
        const float trans[3] = {P[3], P[7], P[11]};
        float qw, qv[3], R[3][3];
        mat_to_quat(m, &qw, qv);                                        /* kernel.cu:125 */
        quat_to_mat(qw, qv, R);                                         /* kernel.cu:126 */
        const float *lim = sdf_limits + 10 * o;
        const int d0 = (int)lim[6], d1 = (int)lim[7], d2 = (int)lim[8]; /* kernel.cu:137-139 */
        const i3 dim = {d0, d1, d2};
        const float delta = lim[9];
        const float *grid = sdf_grids + (size_t)o * d0 * d1 * d2;
        const float eps = epsilons[o], pad = padding_scales[o], clr = clearances[o];
        for (int n = 0; n < num_points; ++n) {
            const f3 pt = {points[3 * n], points[3 * n + 1], points[3 * n + 2]};
            /* so3 * p + t  (so3.hpp:298-300, se3.hpp operator*) */
            f3 uv = cross_fma(qv, pt);
            uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
            const f3 c2 = cross_fma(qv, uv);
            f3 u;
            u.x = (fmaf(qw, uv.x, pt.x) + c2.x) + trans[0];
            u.y = (fmaf(qw, uv.y, pt.y) + c2.y) + trans[1];
            u.z = (fmaf(qw, uv.z, pt.z) + c2.z) + trans[2];
            f3 pg;                                                      /* kernel.cu:140-142 */
            pg.x = (u.x - lim[0]) / (lim[3] - lim[0]) * (float)d0;
            pg.y = (u.y - lim[1]) / (lim[4] - lim[1]) * (float)d1;
            pg.z = (u.z - lim[2]) / (lim[5] - lim[2]) * (float)d2;
            int inb;
            const float value = value_interp(pg, dim, grid, &inb);      /* kernel.cu:147 */
            p_in += inb;
            if (value < clr) collides[n] += 1.0f;                       /* kernel.cu:150-151 */
            float pot, vg[3];
            if (value <= 0) {                                           /* kernel.cu:158-164 */
                const f3 g = grad_interp(pg, dim, grid, delta);
                pot = (float)(-(double)value + 0.5 * (double)eps);
                vg[0] = -g.x; vg[1] = -g.y; vg[2] = -g.z;
            } else if (value > 0 && value <= eps) {                     /* kernel.cu:165-171 */
                const f3 g = grad_interp(pg, dim, grid, delta);
                const float d = value - eps;
                pot = 1 / (2 * eps) * d * d * pad;
                vg[0] = 1 / eps * g.x * d * pad;
                vg[1] = 1 / eps * g.y * d * pad;
                vg[2] = 1 / eps * g.z * d * pad;
            } else {
                continue;                                               /* kernel.cu:172-173 */
            }
            /* rotationMatrix.transpose() * vgrad   kernel.cu:176: coefficient i = (row_i . v).sum() = c0 + (c1 + c2)
             * (Eigen's unrolled tree redux), fused by nvcc as fma(R0i, v0, fma(R1i, v1, R2i*v2)) */
            const float gx = fmaf(R[0][0], vg[0], fmaf(R[1][0], vg[1], R[2][0] * vg[2]));
            const float gy = fmaf(R[0][1], vg[0], fmaf(R[1][1], vg[1], R[2][1] * vg[2]));
            const float gz = fmaf(R[0][2], vg[0], fmaf(R[1][2], vg[1], R[2][2] * vg[2]));
            potentials[n] += pot;                                       /* kernel.cu:186-195, 250-258 */
            potential_grads[3 * n + 0] += gx;
            potential_grads[3 * n + 1] += gy;
            potential_grads[3 * n + 2] += gz;
        }
    }
    return p_in;
}

/* The two interpolation helpers on their own (grid coordinates in, value / gradient out), for the pin against the
 * reference's own getValueInterpolated / getGradientInterpolated (tests/test_oracle_sdf.py). */
void omg_oracle_value_interp(const float *pgrid, int n, const float *grid, int d0, int d1, int d2, float *out) {
    const i3 dim = {d0, d1, d2};
    for (int i = 0; i < n; ++i) {
        const f3 p = {pgrid[3 * i], pgrid[3 * i + 1], pgrid[3 * i + 2]};
        out[i] = value_interp(p, dim, grid, 0);
    }
}

void omg_oracle_grad_interp(const float *pgrid, int n, const float *grid, int d0, int d1, int d2, float delta,
                            float *out) {
    const i3 dim = {d0, d1, d2};
    for (int i = 0; i < n; ++i) {
        const f3 p = {pgrid[3 * i], pgrid[3 * i + 1], pgrid[3 * i + 2]};
        const f3 g = grad_interp(p, dim, grid, delta);
        out[3 * i] = g.x; out[3 * i + 1] = g.y; out[3 * i + 2] = g.z;
    }
}
