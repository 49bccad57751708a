/*
 * TEST INFRASTRUCTURE -- CPU restatement (plain C) of the inverse kinematics the reference calls for goal-set
 * construction: robot_kinematics.inverse_kinematics (ycb_render/robotPose/robot_pykdl.py:257-289) ->
 * KDL::ChainIkSolverPos_NR_JL::CartToJnt with ChainIkSolverVel_pinv and ChainFkSolverPos_recursive, all at their
 * default parameters (robot_pykdl.py:140-146).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may use this file.  Files cited are under orocos_kinematics_dynamics/orocos_kdl/src of the reference tree:
 *
 *   position loop       chainiksolverpos_nr_jl.cpp:61-101   (<= 100 Newton steps, stop when every twist component is
 *                                                            within 1e-6, joints clamped to the limits after each step)
 *   pose error          frames.inl:1133-1143 (diff), frames.cpp:337-431 (Rotation::GetRot / GetRotAngle),
 *                       frames.cpp:118-157 (Vector::Norm / Normalize)
 *   forward kinematics  chainfksolverpos_recursive.cpp, segment.cpp (pose = joint.pose(q) * f_tip),
 *                       joint.cpp:59-85 (RotAxis joint), frames.cpp:304-331 (Rotation::Rot2)
 *   Jacobian            chainjnttojacsolver.cpp:49-95, jacobian.cpp:84-96 (changeRefPoint), frames.inl (Twist::RefPoint)
 *   velocity step       chainiksolvervel_pinv.cpp:61-123 (truncated pseudo-inverse, singular values < 1e-5 dropped)
 *   SVD                 utilities/svd_HH.cpp:56-273 (Householder bidiagonalisation + implicit-shift QR, <= 150 sweeps;
 *                       `anorm` passes through two bool variables there (:43,:136-138), so it is 0 or 1 -- kept)
 *   target orientation  frames.cpp:191-198 (Rotation::Quaternion, not normalised)
 *
 * PARITY STATUS: pinned.  oracle/kdl_ref/Makefile compiles the reference's own KDL sources (Eigen containers replaced
 * by oracle/kdl_ref/eigen_shim) into oracle/_ref/libkdl_ik.so; tests/test_oracle_ik.py requires bit-identical joint
 * solutions and status codes from both on seeded problems, and tests/golden/ik_kdl.npz holds outputs of that library
 * for the GPU box.
 */
#include <math.h>
#include <string.h>

#define NJ 7        /* arm joints */
#define NSEG 8      /* 7 revolute segments + the fixed hand segment */

typedef struct { double M[9]; double p[3]; } frame_t;

static void rot_mul(const double *a, const double *b, double *c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
static void rot_vec(const double *a, const double *v, double *o) {
    for (int i = 0; i < 3; ++i) o[i] = a[3 * i] * v[0] + a[3 * i + 1] * v[1] + a[3 * i + 2] * v[2];
}
static void frame_mul(const frame_t *a, const frame_t *b, frame_t *c) {   /* (M1 M2, M1 p2 + p1) */
    frame_t r;
    double t[3];
    rot_mul(a->M, b->M, r.M);
    rot_vec(a->M, b->p, t);
    for (int i = 0; i < 3; ++i) r.p[i] = t[i] + a->p[i];
    *c = r;
}
static double sqr(double x) { return x * x; }

static double vec_norm(const double *d) {   /* frames.cpp:118-143 */
    double t1 = fabs(d[0]), t2 = fabs(d[1]);
    if (t1 >= t2) {
        t2 = fabs(d[2]);
        if (t1 >= t2) {
            if (t1 == 0) return 0;
            return t1 * sqrt(1 + sqr(d[1] / d[0]) + sqr(d[2] / d[0]));
        }
        return t2 * sqrt(1 + sqr(d[0] / d[2]) + sqr(d[1] / d[2]));
    }
    t1 = fabs(d[2]);
    if (t2 > t1) return t2 * sqrt(1 + sqr(d[0] / d[1]) + sqr(d[2] / d[1]));
    return t1 * sqrt(1 + sqr(d[0] / d[2]) + sqr(d[1] / d[2]));
}

static void rot2(const double *v, double angle, double *R) {   /* frames.cpp:304-331 */
    const double ct = cos(angle), st = sin(angle), vt = 1 - ct;
    const double m_vt_0 = vt * v[0], m_vt_1 = vt * v[1], m_vt_2 = vt * v[2];
    const double m_st_0 = v[0] * st, m_st_1 = v[1] * st, m_st_2 = v[2] * st;
    const double m_vt_0_1 = m_vt_0 * v[1], m_vt_0_2 = m_vt_0 * v[2], m_vt_1_2 = m_vt_1 * v[2];
    R[0] = ct + m_vt_0 * v[0];   R[1] = -m_st_2 + m_vt_0_1;   R[2] = m_st_1 + m_vt_0_2;
    R[3] = m_st_2 + m_vt_0_1;    R[4] = ct + m_vt_1 * v[1];   R[5] = -m_st_0 + m_vt_1_2;
    R[6] = -m_st_1 + m_vt_0_2;   R[7] = m_st_0 + m_vt_1_2;    R[8] = ct + m_vt_2 * v[2];
}

/* the chain as kdl_parser builds it from the URDF: segment s has a revolute joint about `axis` (parent frame,
 * normalised by Joint's constructor) at `origin`, and the tip frame (R_pj, 0) relative to the joint */
typedef struct {
    double axis[NSEG][3], origin[NSEG][3], tipM[NSEG][9];
    int movable[NSEG];
    double qmin[NJ], qmax[NJ];
} chain_t;

static void chain_init(chain_t *c, const double *frames, const double *qmin, const double *qmax) {
    for (int s = 0; s < NSEG; ++s) {
        const double *m = frames + 16 * s;
        const double R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
        const double z[3] = {0.0, 0.0, 1.0};
        double a[3];
        rot_vec(R, z, a);                                   /* f.M * axis */
        const double nrm = vec_norm(a);
        for (int i = 0; i < 3; ++i) { c->axis[s][i] = a[i] / nrm; c->origin[s][i] = m[4 * i + 3]; }
        memcpy(c->tipM[s], R, sizeof(R));
        c->movable[s] = s < NJ;
    }
    memcpy(c->qmin, qmin, sizeof(c->qmin));
    memcpy(c->qmax, qmax, sizeof(c->qmax));
}

static void segment_pose(const chain_t *c, int s, double q, frame_t *out) {
    frame_t joint, tip;
    if (c->movable[s]) rot2(c->axis[s], q, joint.M);
    else { memset(joint.M, 0, sizeof(joint.M)); joint.M[0] = joint.M[4] = joint.M[8] = 1.0; }
    for (int i = 0; i < 3; ++i) { joint.p[i] = c->movable[s] ? c->origin[s][i] : 0.0; tip.p[i] = 0.0; }
    memcpy(tip.M, c->tipM[s], sizeof(tip.M));
    if (!c->movable[s])                                     /* Joint::None: pose = identity * f_tip, f_tip = the frame */
        for (int i = 0; i < 3; ++i) tip.p[i] = c->origin[s][i];
    frame_mul(&joint, &tip, out);
}

static void identity(frame_t *f) {
    memset(f, 0, sizeof(*f));
    f->M[0] = f->M[4] = f->M[8] = 1.0;
}

static void chain_fk(const chain_t *c, const double *q, frame_t *out) {
    frame_t T, P;
    identity(&T);
    int j = 0;
    for (int s = 0; s < NSEG; ++s) {
        segment_pose(c, s, c->movable[s] ? q[j] : 0.0, &P);
        if (c->movable[s]) ++j;
        frame_mul(&T, &P, &T);
    }
    *out = T;
}

static void cross(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* chainjnttojacsolver.cpp:49-95: J[6][NJ], rows 0-2 linear, 3-5 angular, reference point = chain tip */
static void chain_jacobian(const chain_t *c, const double *q, double J[6][NJ]) {
    frame_t T, P, total;
    double tw_rot[3] = {0, 0, 0}, tw_vel[3] = {0, 0, 0};
    identity(&T);
    memset(J, 0, sizeof(double) * 6 * NJ);
    int j = 0, k = 0;
    for (int s = 0; s < NSEG; ++s) {
        if (c->movable[s]) {
            segment_pose(c, s, q[j], &P);
            frame_mul(&T, &P, &total);
            /* Segment::twist(q, 1): joint twist (0, axis) moved to the tip, which coincides with the joint origin */
            double jr[3], jv[3], zero[3] = {0, 0, 0}, cr[3], jM[9], tipp[3];
            rot2(c->axis[s], q[j], jM);
            rot_vec(jM, zero, tipp);                          /* joint.pose(q).M * f_tip.p, f_tip.p = 0 */
            for (int i = 0; i < 3; ++i) jr[i] = c->axis[s][i] * 1.0;
            cross(jr, tipp, cr);
            for (int i = 0; i < 3; ++i) jv[i] = 0.0 + cr[i];
            rot_vec(T.M, jv, tw_vel);
            rot_vec(T.M, jr, tw_rot);
        } else {
            segment_pose(c, s, 0.0, &P);
            frame_mul(&T, &P, &total);
        }
        double d[3];
        for (int i = 0; i < 3; ++i) d[i] = total.p[i] - T.p[i];
        for (int col = 0; col < NJ; ++col) {                  /* changeRefPoint: vel += rot x d */
            double r[3] = {J[3][col], J[4][col], J[5][col]}, cr[3];
            cross(r, d, cr);
            for (int i = 0; i < 3; ++i) J[i][col] = J[i][col] + cr[i];
        }
        if (c->movable[s]) {
            for (int i = 0; i < 3; ++i) { J[i][k] = tw_vel[i]; J[3 + i][k] = tw_rot[i]; }
            ++k; ++j;
        }
        T = total;
    }
}

static double pythag(double a, double b) {   /* svd_HH.cpp:29-44 */
    const double at = fabs(a), bt = fabs(b);
    if (at > bt) { const double ct = bt / at; return at * sqrt(1.0 + ct * ct); }
    if (bt == 0) return 0.0;
    { const double ct = at / bt; return bt * sqrt(1.0 + ct * ct); }
}
static double sign_of(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

/* svd_HH.cpp:56-273 for a 6 x 7 matrix: U [6][7], w [7], V [7][7]; returns 0, or -2 like the original */
static int svd_hh(const double A[6][NJ], double U[6][NJ], double *w, double V[NJ][NJ], int maxiter) {
    const int rows = 6, cols = NJ;
    double tmp[NJ];
    int i, its = -1, j, jj, k, nm = 0, ppi = 0, flag;
    double anorm = 0, c = 0, f = 0, h = 0, s = 0, scale = 0, x = 0, y = 0, z = 0, g = 0;
    for (i = 0; i < rows; i++) for (j = 0; j < cols; j++) U[i][j] = A[i][j];
    for (i = 0; i < cols; i++) {                                        /* Householder, bidiagonal form */
        ppi = i + 1;
        tmp[i] = scale * g;
        g = s = scale = 0.0;
        if (i < rows) {
            for (k = i; k < rows; k++) scale += fabs(U[k][i]);
            if (scale) {
                for (k = i; k < rows; k++) { U[k][i] /= scale; s += U[k][i] * U[k][i]; }
                f = U[i][i];
                g = -sign_of(sqrt(s), f);
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
            if (scale) {
                for (k = ppi; k < cols; k++) { U[i][k] /= scale; s += U[i][k] * U[i][k]; }
                f = U[i][ppi];
                g = -sign_of(sqrt(s), f);
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
        {   /* (sic) the original stores both operands of the max in bool variables */
            const int m1 = anorm != 0.0, m2 = (fabs(w[i]) + fabs(tmp[i])) != 0.0;
            anorm = m1 > m2 ? (double)m1 : (double)m2;
        }
    }
    for (i = cols - 1; i >= 0; i--) {                                   /* right-hand transformations */
        if (i < cols - 1) {
            if (g) {
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
    for (i = (cols - 1 < rows - 1 ? cols - 1 : rows - 1); i >= 0; i--) { /* left-hand transformations */
        ppi = i + 1;
        g = w[i];
        for (j = ppi; j < cols; j++) U[i][j] = 0.0;
        if (g) {
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
        ++U[i][i];
    }
    for (k = cols - 1; k >= 0; k--) {                                   /* diagonalisation */
        for (its = 1; its <= maxiter; its++) {
            flag = 1;
            for (ppi = k; ppi >= 0; ppi--) {
                nm = ppi - 1;
                if ((fabs(tmp[ppi]) + anorm) == anorm) { flag = 0; break; }
                if ((fabs(w[nm] + anorm) == anorm)) break;
            }
            if (flag) {
                c = 0.0;
                s = 1.0;
                for (i = ppi; i <= k; i++) {
                    f = s * tmp[i];
                    tmp[i] = c * tmp[i];
                    if ((fabs(f) + anorm) == anorm) break;
                    g = w[i];
                    h = pythag(f, g);
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
            g = pythag(f, 1.0);
            f = ((x - z) * (x + z) + h * ((y / (f + sign_of(g, f))) - h)) / x;
            c = s = 1.0;
            for (j = ppi; j <= nm; j++) {
                i = j + 1;
                g = tmp[i];
                y = w[i];
                h = s * g;
                g = c * g;
                z = pythag(f, h);
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
                z = pythag(f, h);
                w[j] = z;
                if (z) { z = 1.0 / z; c = f * z; s = h * z; }
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

/* frames.cpp:359-431 with eps = KDL::epsilon = 1e-6, then axis * angle (frames.cpp:337-345) */
static void get_rot(const double *d, double *out) {
    const double eps = 0.000001, eps2 = eps * 10;
    double x, y, z, angle;
    if ((fabs(d[1] - d[3]) < eps) && (fabs(d[2] - d[6]) < eps) && (fabs(d[5] - d[7]) < eps)) {
        if ((fabs(d[1] + d[3]) < eps2) && (fabs(d[2] + d[6]) < eps2) && (fabs(d[5] + d[7]) < eps2) &&
            (fabs(d[0] + d[4] + d[8] - 3) < eps2)) {
            out[0] = 0 * 0.0; out[1] = 0 * 0.0; out[2] = 1 * 0.0;
            return;
        }
        angle = M_PI;
        const double xx = (d[0] + 1) / 2, yy = (d[4] + 1) / 2, zz = (d[8] + 1) / 2;
        const double xy = (d[1] + d[3]) / 4, xz = (d[2] + d[6]) / 4, yz = (d[5] + d[7]) / 4;
        if ((xx > yy) && (xx > zz)) { x = sqrt(xx); y = xy / x; z = xz / x; }
        else if (yy > zz) { y = sqrt(yy); x = xy / y; z = yz / y; }
        else { z = sqrt(zz); x = xz / z; y = yz / z; }
        out[0] = x * angle; out[1] = y * angle; out[2] = z * angle;
        return;
    }
    const double f = (d[0] + d[4] + d[8] - 1) / 2;
    angle = acos(fmax(-1.0, fmin(1.0, f)));
    double a[3] = {d[7] - d[5], d[2] - d[6], d[3] - d[1]};
    const double v = vec_norm(a);
    if (v < eps) { a[0] = 1; a[1] = 0; a[2] = 0; }
    else { a[0] = a[0] / v; a[1] = a[1] / v; a[2] = a[2] / v; }
    out[0] = a[0] * angle; out[1] = a[1] * angle; out[2] = a[2] * angle;
}

/* One inverse_kinematics call.  frames: [8][16] row-major 4x4 parent->joint transforms (robot_kinematics._pose_0[:8]);
 * returns KDL's status: 0 found, -5 iteration limit, -100 SVD failure; *iterations = Newton steps taken. */
int omg_oracle_ik(const double *frames, const double *qmin, const double *qmax, const double *position,
                  const double *quat_xyzw, const double *seed, double *result, int *iterations) {
    chain_t c;
    chain_init(&c, frames, qmin, qmax);
    frame_t goal, f;
    {
        const double x = quat_xyzw[0], y = quat_xyzw[1], z = quat_xyzw[2], w = quat_xyzw[3];
        const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
        const double R[9] = {w2 + x2 - y2 - z2, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y,
                             2 * x * y + 2 * w * z, w2 - x2 + y2 - z2, 2 * y * z - 2 * w * x,
                             2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, w2 - x2 - y2 + z2};
        memcpy(goal.M, R, sizeof(R));
        memcpy(goal.p, position, sizeof(goal.p));
    }
    double q[NJ];
    memcpy(q, seed, sizeof(q));
    const unsigned maxiter = 100;
    const double eps = 1e-6, svd_eps = 0.00001;
    unsigned i;
    int status = 0;
    for (i = 0; i < maxiter; i++) {
        chain_fk(&c, q, &f);
        double tw[6], Mi[9], Rrel[9], rv[3], rr[3];
        for (int a = 0; a < 3; ++a) tw[a] = (goal.p[a] - f.p[a]) / 1.0;
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Mi[3 * a + b] = f.M[3 * b + a];
        rot_mul(Mi, goal.M, Rrel);
        get_rot(Rrel, rv);
        rot_vec(f.M, rv, rr);
        for (int a = 0; a < 3; ++a) tw[3 + a] = rr[a] / 1.0;
        int zero = 1;
        for (int a = 0; a < 6; ++a) { const double t = tw[a] - 0.0; if (!((eps > t) && (t > -eps))) zero = 0; }
        if (zero) break;
        double J[6][NJ], U[6][NJ], S[NJ], V[NJ][NJ], tmp[NJ], dq[NJ];
        chain_jacobian(&c, q, J);
        if (svd_hh(J, U, S, V, 150) != 0) { status = -100; break; }     /* E_IKSOLVERVEL_FAILED */
        for (int a = 0; a < NJ; ++a) {
            double sum = 0.0;
            for (int b = 0; b < 6; ++b) sum += U[b][a] * tw[b];
            tmp[a] = fabs(S[a]) < svd_eps ? 0.0 : sum / S[a];
        }
        for (int a = 0; a < NJ; ++a) {
            double sum = 0.0;
            for (int b = 0; b < NJ; ++b) sum += V[a][b] * tmp[b];
            dq[a] = sum;
        }
        for (int a = 0; a < NJ; ++a) q[a] = q[a] + dq[a];
        for (int a = 0; a < NJ; ++a) if (q[a] < c.qmin[a]) q[a] = c.qmin[a];
        for (int a = 0; a < NJ; ++a) if (q[a] > c.qmax[a]) q[a] = c.qmax[a];
    }
    memcpy(result, q, sizeof(q));
    if (iterations) *iterations = (int)i;
    if (status) return status;
    return i != maxiter ? 0 : -5;
}

/* The chain of solves of solve_one_pose_ik (omg/planner.py:38-87) for one (grasp, seed): targets [T][7] =
 * (position, quaternion xyzw), each seeded with the previous solution; stops at the first failure.
 * Returns the number of successful solves; sols [T][7]. */
int omg_oracle_ik_chain(const double *frames, const double *qmin, const double *qmax, const double *targets, int T,
                        const double *seed, double *sols) {
    double q[NJ];
    memcpy(q, seed, sizeof(q));
    for (int t = 0; t < T; ++t) {
        double r[NJ];
        if (omg_oracle_ik(frames, qmin, qmax, targets + 7 * t, targets + 7 * t + 3, q, r, 0) < 0) return t;
        memcpy(sols + NJ * t, r, sizeof(r));
        memcpy(q, r, sizeof(q));
    }
    return T;
}

void omg_oracle_fk_hand(const double *frames, const double *q, double *out16) {
    chain_t c;
    double lim[NJ] = {0};
    chain_init(&c, frames, lim, lim);
    frame_t f;
    chain_fk(&c, q, &f);
    for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) out16[4 * r + k] = f.M[3 * r + k]; out16[4 * r + 3] = f.p[r]; }
    out16[12] = out16[13] = out16[14] = 0.0; out16[15] = 1.0;
}
