// Fused CHOMP iteration for sm_100a: one CTA owns one trajectory for the whole iteration.
//
//   phase 0  stage xi/start/end/goal rows, object records, body points in shared memory
//   phase 1  Panda forward kinematics for the n waypoints + start + end in fp64 registers
//            (robot_pykdl.py:148-215), link frames + joint axes/"origins" to shared memory
//   phase 1b bounding-sphere cull of every (waypoint, link) against every object's grid box
//   phase 2  half-warp per (waypoint, link), lane per body point: fp64 point placement, fp32 SDF sampling
//            (layers/sdf_matching_loss_kernel.cu:97-181), warp-shuffle argmax / reductions over body points
//   phase 3  top-k membership threshold (radix select) when more than k points have potential
//   phase 4  CHOMP functional gradient + Jacobian pull-back (omg/cost.py:24-43, 92-110, 362-423)
//   phase 5  smoothness term, clip, weights, norms (omg/cost.py:425-532)
//   phase 6  covariant update  -eta*Ainv*g (+ goal-set projection)  (omg/optimizer.py:88-135, core.py:43-51)
//   phase 7  smooth joint-limit projection (omg/optimizer.py:148-164)
//   phase 8  write xi and the info row
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/omgb200.h"
#include "sdf_device.cuh"

namespace omgb {

constexpr int NL = OMGB_NUM_LINKS;
constexpr int ND = OMGB_NUM_DOF;
constexpr int NJ = 9;   // joint-info slots per waypoint: 7 arm joints, finger joint 8, finger joint 9
constexpr int NS = 8;   // gradient slots per link (<= 7 arm ancestors + own prismatic joint)

// Robot constants resident in HBM (read through L1; every access is warp-uniform or staged to smem).
struct RobotConst {
    double P0[10][12];     // pose_0[i]: rotation row-major [0:9], translation [9:12]
    double CO[10][12];     // center_offset[j]
    double ja[10][3];      // joint axis in the link frame: tip2joint_R * axis
    double jo[10][3];      // joint "origin" in the link frame: tip2joint_R * origin + tip2joint_t
    double pts[10][OMGB_MAX_BODY_POINTS][3];
    float sph[10][4];      // bounding sphere of the link's body points (link frame centre, radius)
    double lower[ND], upper[ND];
    int p;                 // body points per link
    int pad_;
};

struct StepArgs {
    const ObjRec *objs;
    const float *grids;
    const RobotConst *robot;
    const double *Ainv;      // [n,n]
    const double *proj;      // [n,c]
    double *xi;              // [B,n,9]
    const double *start;     // [B,9]
    const double *end;       // [B,9]
    const double *goal_rows; // [B,c,9]
    const uint8_t *active;   // [B] or null
    uint8_t *done;           // [B] or null (plan mode)
    double *grad_out;        // [B,n,9] or null
    double *info;            // [B,16]
    float *dbg_pot;          // [B,n,10,p] or null
    float *dbg_pts;          // [B,n,10,p,3] or null
    double *row_obs;         // [B,n] or null: obstacle cost per waypoint row (obs_cost.sum(-1)); zeroed by the caller
    int num_objects;
    int batch;
    int iteration;           // index inside a plan (for the t > 0 rule of planner.py:627)
    int stop_on_terminate;
    omgb_step_params_t prm;
};

// ----------------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; every thread gets the result.  scratch: >= 33 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

__device__ __forceinline__ void xform(const double *F, double px, double py, double pz, double &x, double &y,
                                      double &z) {
    x = fma(F[0], px, fma(F[1], py, fma(F[2], pz, F[9])));
    y = fma(F[3], px, fma(F[4], py, fma(F[5], pz, F[10])));
    z = fma(F[6], px, fma(F[7], py, fma(F[8], pz, F[11])));
}

// C = A * B for rigid transforms stored as R[9] row-major + t[3].
__device__ __forceinline__ void compose(const double *A, const double *B, double *C) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            C[3 * r + c] = fma(A[3 * r + 0], B[c], fma(A[3 * r + 1], B[3 + c], A[3 * r + 2] * B[6 + c]));
        C[9 + r] = fma(A[3 * r + 0], B[9], fma(A[3 * r + 1], B[10], fma(A[3 * r + 2], B[11], A[9 + r])));
    }
}

// Forward kinematics of one configuration q[9] (rad): 10 body-point frames (T_j * center_offset_j) and,
// when jinfo != null, joint axis + reference "origin" for the 7 arm joints and the 2 finger axes.
// robot_pykdl.py:148-215; the rotX(+-pi)/column-flip pair of :166,174-176 cancels and is omitted.
__device__ void panda_fk(const RobotConst *__restrict__ rc, const double *q, double *frames, double *jinfo) {
    double T[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    double N[12];
#pragma unroll 1
    for (int i = 0; i < 7; ++i) {
        double B[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) B[k] = rc->P0[i][k];
        double s, c;
        sincos(q[i], &s, &c);
        compose(T, B, N);   // T * pose_0[i]
        // ... * Rz(q_i): rotate the first two columns
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double a = N[3 * r], b = N[3 * r + 1];
            T[3 * r] = fma(a, c, b * s);
            T[3 * r + 1] = fma(b, c, -(a * s));
            T[3 * r + 2] = N[3 * r + 2];
            T[9 + r] = N[9 + r];
        }
        double C[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) C[k] = rc->CO[i][k];
        compose(T, C, N);
#pragma unroll
        for (int k = 0; k < 12; ++k) frames[12 * i + k] = N[k];
        if (jinfo) {
            const double ax = rc->ja[i][0], ay = rc->ja[i][1], az = rc->ja[i][2];
            const double ox = rc->jo[i][0], oy = rc->jo[i][1], oz = rc->jo[i][2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                jinfo[6 * i + r] = fma(T[3 * r], ax, fma(T[3 * r + 1], ay, T[3 * r + 2] * az));
                jinfo[6 * i + 3 + r] = fma(T[3 * r], ox, fma(T[3 * r + 1], oy, fma(T[3 * r + 2], oz, T[9 + r])));
            }
        }
    }
    double H[12], B[12], C[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) B[k] = rc->P0[7][k];
    compose(T, B, H);   // hand
#pragma unroll
    for (int k = 0; k < 12; ++k) C[k] = rc->CO[7][k];
    compose(H, C, N);
#pragma unroll
    for (int k = 0; k < 12; ++k) frames[12 * 7 + k] = N[k];
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
#pragma unroll
        for (int k = 0; k < 12; ++k) B[k] = rc->P0[8 + f][k];
        B[10] += (f == 0) ? q[7] : -q[8];   // robot_pykdl.py:181-184
        compose(H, B, T);
#pragma unroll
        for (int k = 0; k < 12; ++k) C[k] = rc->CO[8 + f][k];
        compose(T, C, N);
#pragma unroll
        for (int k = 0; k < 12; ++k) frames[12 * (8 + f) + k] = N[k];
        if (jinfo) {
            const double ax = rc->ja[8 + f][0], ay = rc->ja[8 + f][1], az = rc->ja[8 + f][2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                jinfo[6 * (7 + f) + r] = fma(T[3 * r], ax, fma(T[3 * r + 1], ay, T[3 * r + 2] * az));
                jinfo[6 * (7 + f) + 3 + r] = 0.0;
            }
        }
    }
}

// number of gradient slots of link j and the joint-info slot / DOF column of slot s
__device__ __forceinline__ int link_slots(int j) { return j < 7 ? j + 1 : (j == 7 ? 7 : 8); }

// CHOMP functional gradient of one body point (omg/cost.py:24-43) pulled back through the point Jacobian
// (omg/cost.py:92-110).  x, xp, xn: the point at waypoint i, i-1, i+1.  Writes g[0..slots) and returns
// c * |v| (the point's obstacle cost).
__device__ __forceinline__ double functional_grad(const double *jinfo_i, int j, double x, double y, double z,
                                                  double xpx, double xpy, double xpz, double xnx, double xny,
                                                  double xnz, double c, double gcx, double gcy, double gcz,
                                                  double dt, double *g) {
    const double idt = 1.0 / dt;
    const double vx = (x - xpx) * idt, vy = (y - xpy) * idt, vz = (z - xpz) * idt;
    const double idt2 = idt * idt;
    const double ax = (xpx - 2.0 * x + xnx) * idt2, ay = (xpy - 2.0 * y + xny) * idt2,
                 az = (xpz - 2.0 * z + xnz) * idt2;
    const double speed = sqrt(vx * vx + vy * vy + vz * vz);
    const double inv = 1.0 / (speed + 1e-8);
    const double hx = vx * inv, hy = vy * inv, hz = vz * inv;
    const double ha = hx * ax + hy * ay + hz * az;
    const double hg = hx * gcx + hy * gcy + hz * gcz;
    const double ks = c / (speed * speed + 1e-8);
    const double wx = speed * (gcx - hx * hg) - ks * (ax - hx * ha);
    const double wy = speed * (gcy - hy * hg) - ks * (ay - hy * ha);
    const double wz = speed * (gcz - hz * hg) - ks * (az - hz * ha);
    const int ns = link_slots(j);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double val = 0.0;
        if (s < ns) {
            if (s < 7) {
                const double *ji = jinfo_i + 6 * s;
                const double rx = x - ji[3], ry = y - ji[4], rz = z - ji[5];
                const double jx = ji[1] * rz - ji[2] * ry, jy = ji[2] * rx - ji[0] * rz,
                             jz = ji[0] * ry - ji[1] * rx;
                val = jx * wx + jy * wy + jz * wz;
            } else {   // prismatic finger joint: the column is the axis itself (cost.py:106-108)
                const double *ji = jinfo_i + 6 * (j - 1);   // link 8 -> slot 7, link 9 -> slot 8
                val = ji[0] * wx + ji[1] * wy + ji[2] * wz;
            }
        }
        g[s] = val;
    }
    return c * speed;
}

struct SmemLayout {
    int n, c, lpi, nobj, p;
    size_t off_xi, off_start, off_end, off_goal, off_frames, off_jinfo, off_lg, off_grad, off_u, off_viol,
        off_red, off_pts, off_mask, off_best, off_bestp, off_objs, off_hist, total;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline SmemLayout make_layout(int n, int c, int lpi, int nobj, int p) {
    SmemLayout L;
    L.n = n; L.c = c; L.lpi = lpi; L.nobj = nobj; L.p = p;
    size_t o = 0;
    L.off_xi = o; o += sizeof(double) * n * ND;
    L.off_start = o; o += sizeof(double) * ND;
    L.off_end = o; o += sizeof(double) * ND;
    L.off_goal = o; o += sizeof(double) * (c > 0 ? c : 1) * ND;
    L.off_frames = o; o += sizeof(double) * (n + 2) * NL * 12;
    L.off_jinfo = o; o += sizeof(double) * n * NJ * 6;
    // link gradients [n*10][8] fp64; aliased with the fp32 potential array [n*10][lpi] of the top-k path
    size_t lg = sizeof(double) * n * NL * NS, pot = sizeof(float) * n * NL * lpi;
    L.off_lg = o; o += (lg > pot ? lg : pot);
    L.off_grad = o; o += sizeof(double) * n * ND;
    L.off_u = o; o += sizeof(double) * n * ND;
    L.off_viol = o; o += sizeof(double) * n * ND;
    L.off_red = o; o += sizeof(double) * 40;
    L.off_pts = o; o += sizeof(double) * NL * p * 3;
    L.off_mask = o; o += sizeof(unsigned long long) * n * NL;
    L.off_best = o; o += sizeof(float) * n * NL;
    L.off_bestp = o; o += sizeof(int) * n * NL;
    o = align_up(o, 16);
    L.off_objs = o; o += sizeof(ObjRec) * nobj;
    L.off_hist = o; o += sizeof(int) * 264;
    L.total = align_up(o, 16);
    return L;
}

// ----------------------------------------------------------------------------------------------------
// the fused iteration
// ----------------------------------------------------------------------------------------------------
template <int LPI>
__global__ void __launch_bounds__(512, 1) chomp_step_kernel(const StepArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x;
    if (b >= a.batch) return;
    if (a.active && !a.active[b]) return;
    if (a.done && a.done[b]) return;

    const omgb_step_params_t &prm = a.prm;
    const int n = prm.n_waypoints, c = prm.constraint_rows;
    const RobotConst *__restrict__ rc = a.robot;
    const int P = rc->p;
    const int O = a.num_objects;
    const SmemLayout L = make_layout(n, c, LPI, O, P);
    double *s_xi = reinterpret_cast<double *>(smem + L.off_xi);
    double *s_start = reinterpret_cast<double *>(smem + L.off_start);
    double *s_end = reinterpret_cast<double *>(smem + L.off_end);
    double *s_goal = reinterpret_cast<double *>(smem + L.off_goal);
    double *s_frames = reinterpret_cast<double *>(smem + L.off_frames);
    double *s_jinfo = reinterpret_cast<double *>(smem + L.off_jinfo);
    double *s_lg = reinterpret_cast<double *>(smem + L.off_lg);
    float *s_pot = reinterpret_cast<float *>(smem + L.off_lg);
    double *s_grad = reinterpret_cast<double *>(smem + L.off_grad);
    double *s_u = reinterpret_cast<double *>(smem + L.off_u);
    double *s_viol = reinterpret_cast<double *>(smem + L.off_viol);
    double *s_red = reinterpret_cast<double *>(smem + L.off_red);
    double *s_pts = reinterpret_cast<double *>(smem + L.off_pts);
    unsigned long long *s_mask = reinterpret_cast<unsigned long long *>(smem + L.off_mask);
    float *s_best = reinterpret_cast<float *>(smem + L.off_best);
    int *s_bestp = reinterpret_cast<int *>(smem + L.off_bestp);
    ObjRec *s_objs = reinterpret_cast<ObjRec *>(smem + L.off_objs);
    int *s_hist = reinterpret_cast<int *>(smem + L.off_hist);

    const int tid = threadIdx.x, nthr = blockDim.x;
    const double dt = prm.time_interval;
    const bool topk_mode = prm.top_k_collision > 0;
    const bool goal_set = prm.goal_set_proj != 0;

    // ---- phase 0: stage ---------------------------------------------------------------------------
    double *g_xi = a.xi + (size_t)b * n * ND;
    for (int k = tid; k < n * ND; k += nthr) s_xi[k] = g_xi[k];
    if (tid < ND) {
        s_start[tid] = a.start[(size_t)b * ND + tid];
        s_end[tid] = a.end[(size_t)b * ND + tid];
    }
    for (int k = tid; k < c * ND; k += nthr) s_goal[k] = a.goal_rows[(size_t)b * c * ND + k];
    {
        const int words = (int)(sizeof(ObjRec) / 4) * O;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.objs);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_objs);
        for (int k = tid; k < words; k += nthr) dst[k] = src[k];
    }
    for (int k = tid; k < NL * P * 3; k += nthr) {
        const int j = k / (P * 3), r = k - j * P * 3;
        s_pts[k] = rc->pts[j][r / 3][r % 3];
    }
    __syncthreads();

    // ---- phase 1: forward kinematics (n waypoints, then start, then end) ---------------------------
    for (int cfg = tid; cfg < n + 2; cfg += nthr) {
        const double *q = (cfg < n) ? (s_xi + cfg * ND) : (cfg == n ? s_start : s_end);
        double ql[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) ql[d] = q[d];
        panda_fk(rc, ql, s_frames + (size_t)cfg * NL * 12, cfg < n ? s_jinfo + (size_t)cfg * NJ * 6 : nullptr);
    }
    __syncthreads();

    // ---- phase 1b: sphere cull ---------------------------------------------------------------------
    for (int li = tid; li < n * NL; li += nthr) {
        const int j = li % NL;
        const double *F = s_frames + (size_t)li * 12;
        double cx, cy, cz;
        xform(F, (double)rc->sph[j][0], (double)rc->sph[j][1], (double)rc->sph[j][2], cx, cy, cz);
        const float fx = (float)cx, fy = (float)cy, fz = (float)cz, rad = rc->sph[j][3];
        unsigned long long m = 0ull;
        for (int o = 0; o < O; ++o) {
            const ObjRec &ob = s_objs[o];
            if (ob.dis > 0.0f) continue;
            const float qx = ob.r[0] * fx + ob.r[1] * fy + ob.r[2] * fz + ob.tx;
            const float qy = ob.r[3] * fx + ob.r[4] * fy + ob.r[5] * fz + ob.ty;
            const float qz = ob.r[6] * fx + ob.r[7] * fy + ob.r[8] * fz + ob.tz;
            const float s = rad + ob.cull_pad;
            const bool hit = (qx > ob.lox - s) & (qx < ob.hix + s) & (qy > ob.loy - s) & (qy < ob.hiy + s) &
                             (qz > ob.loz - s) & (qz < ob.hiz + s);
            if (hit) m |= (1ull << o);
        }
        s_mask[li] = m;
    }
    __syncthreads();

    // ---- phase 2: body points x objects ------------------------------------------------------------
    constexpr int GPW = 32 / LPI;                 // link instances per warp
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int sub = lane / LPI, pl = lane % LPI;  // which instance of the warp, which body point
    const bool finger_soft = (prm.uncheck_finger_collision == -1);
    int t_nnz = 0, t_pin = 0, t_col = 0;
    double t_cost = 0.0;
    const int n_li = n * NL;
    for (int base = warp * GPW; base < n_li; base += nwarps * GPW) {
        const int li = base + sub;
        const bool live = (li < n_li) && (pl < P);
        const int lic = li < n_li ? li : n_li - 1;
        const int i = lic / NL, j = lic - i * NL;
        const double *F = s_frames + (size_t)lic * 12;
        const double *bp = s_pts + ((size_t)j * P + (pl < P ? pl : 0)) * 3;
        double X, Y, Z;
        xform(F, bp[0], bp[1], bp[2], X, Y, Z);
        const float x = (float)X, y = (float)Y, z = (float)Z;   // omg/cost.py:136 .float()
        unsigned long long m = live ? s_mask[lic] : 0ull;
        float pot = 0.0f, col = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
        while (m) {
            const int o = __ffsll((long long)m) - 1;
            m &= m - 1;
            float po, co;
            bool inb;
            if (topk_mode) {
                inb = pair_potential(s_objs[o], a.grids, x, y, z, po, co);
            } else {
                float ax, ay, az;
                inb = pair_full(s_objs[o], a.grids, x, y, z, po, ax, ay, az, co);
                gx = __fadd_rn(gx, ax); gy = __fadd_rn(gy, ay); gz = __fadd_rn(gz, az);
            }
            pot = __fadd_rn(pot, po);
            col = __fadd_rn(col, co);
            t_pin += inb ? 1 : 0;
        }
        if (finger_soft && j >= 8) {   // omg/cost.py:350-353
            pot = __fmul_rn(pot, 0.1f); gx = __fmul_rn(gx, 0.1f); gy = __fmul_rn(gy, 0.1f);
            gz = __fmul_rn(gz, 0.1f); col = 0.0f;
        }
        if (live) {
            t_nnz += (pot > 0.0f) ? 1 : 0;
            t_col += (int)col;
            if (a.dbg_pot) a.dbg_pot[((size_t)b * n_li + li) * P + pl] = pot;
            if (a.dbg_pts) {
                float *d = a.dbg_pts + (((size_t)b * n_li + li) * P + pl) * 3;
                d[0] = x; d[1] = y; d[2] = z;
            }
        }
        if (topk_mode) {
            if (li < n_li) s_pot[(size_t)li * LPI + pl] = live ? pot : 0.0f;
            // argmax over the body points of this link instance; ties -> highest point index
            float bv = live ? pot : -1.0f;
            int bi = pl;
#pragma unroll
            for (int off = LPI / 2; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, off, LPI);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off, LPI);
                if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
            }
            if (li < n_li && pl == 0) { s_best[li] = bv; s_bestp[li] = bi; }
        } else {
            // full-sum mode: functional gradient of every point with non-zero potential, reduced over
            // the link instance's body points with warp shuffles
            double g[NS];
            double cst = 0.0;
#pragma unroll
            for (int s = 0; s < NS; ++s) g[s] = 0.0;
            if (live && (pot != 0.0f || gx != 0.0f || gy != 0.0f || gz != 0.0f)) {
                const double *Fp = (i > 0) ? (F - NL * 12) : (s_frames + ((size_t)n * NL + j) * 12);
                const double *Fn = (i < n - 1) ? (F + NL * 12) : (s_frames + ((size_t)(n + 1) * NL + j) * 12);
                double xp, yp, zp, xn, yn, zn;
                xform(Fp, bp[0], bp[1], bp[2], xp, yp, zp);
                xform(Fn, bp[0], bp[1], bp[2], xn, yn, zn);
                cst = functional_grad(s_jinfo + (size_t)i * NJ * 6, j, X, Y, Z, xp, yp, zp, xn, yn, zn,
                                      (double)pot, (double)gx, (double)gy, (double)gz, dt, g);
            }
#pragma unroll
            for (int off = LPI / 2; off > 0; off >>= 1) {
#pragma unroll
                for (int s = 0; s < NS; ++s) g[s] += __shfl_xor_sync(0xffffffffu, g[s], off, LPI);
                cst += __shfl_xor_sync(0xffffffffu, cst, off, LPI);
            }
            if (li < n_li && pl == 0) {
#pragma unroll
                for (int s = 0; s < NS; ++s) s_lg[(size_t)li * NS + s] = g[s];
                t_cost += cst;
                if (a.row_obs) atomicAdd(a.row_obs + (size_t)b * n + i, cst);
            }
        }
    }
    const int nnz = (int)(block_sum((double)t_nnz, s_red) + 0.5);
    const int p_in = (int)(block_sum((double)t_pin, s_red) + 0.5);
    const int collide = (int)(block_sum((double)t_col, s_red) + 0.5);

    double obs_sum = 0.0;
    if (topk_mode) {
        // ---- phase 3: membership threshold = k-th largest potential (bit pattern; potentials >= 0) ----
        const int K = prm.top_k_collision;
        const int n_slots = n_li * LPI;
        uint32_t tau = 1u;   // "pot >= tau" <=> pot > 0
        if (nnz > K) {
            uint32_t prefix = 0u, pmask = 0u;
            int remaining = K;
            for (int shift = 24; shift >= 0; shift -= 8) {
                for (int k = tid; k < 256; k += nthr) s_hist[k] = 0;
                __syncthreads();
                for (int k = tid; k < n_slots; k += nthr) {
                    const uint32_t u = __float_as_uint(s_pot[k]);
                    if ((u & pmask) == prefix) atomicAdd(&s_hist[(u >> shift) & 255u], 1);
                }
                __syncthreads();
                if (tid == 0) {
                    int acc = 0, bsel = 0;
                    for (int bkt = 255; bkt >= 0; --bkt) {
                        if (acc + s_hist[bkt] >= remaining) { bsel = bkt; break; }
                        acc += s_hist[bkt];
                    }
                    s_hist[256] = bsel;
                    s_hist[257] = remaining - acc;
                }
                __syncthreads();
                prefix |= ((uint32_t)s_hist[256]) << shift;
                pmask |= 255u << shift;
                remaining = s_hist[257];
                __syncthreads();
            }
            tau = prefix;   // ties at tau are all kept (reference: unstable argsort, order undefined)
        }
        // ---- phase 4a: obstacle cost = sum over members of c*|v| (cost.py:30,416), links 0..7 ---------
        const int jmax = prm.consider_finger ? NL : NL - 2;
        double acc = 0.0;
        for (int k = tid; k < n_slots; k += nthr) {
            const float pv = s_pot[k];
            const int li = k / LPI, p = k - li * LPI;
            const int i = li / NL, j = li - i * NL;
            if (p < P && j < jmax && pv > 0.0f && __float_as_uint(pv) >= tau) {
                const double *F = s_frames + (size_t)li * 12;
                const double *Fp = (i > 0) ? (F - NL * 12) : (s_frames + ((size_t)n * NL + j) * 12);
                const double *bp = s_pts + ((size_t)j * P + p) * 3;
                double X, Y, Z, xp, yp, zp;
                xform(F, bp[0], bp[1], bp[2], X, Y, Z);
                xform(Fp, bp[0], bp[1], bp[2], xp, yp, zp);
                const double vx = (X - xp) / dt, vy = (Y - yp) / dt, vz = (Z - zp) / dt;
                acc += (double)pv * sqrt(vx * vx + vy * vy + vz * vz);
            }
        }
        obs_sum = block_sum(acc, s_red) * (double)n;   // added to every waypoint row (SURVEY A-3)
        __syncthreads();   // s_pot is dead from here on; s_lg (same storage) is written next
        // ---- phase 4b: one winner per (waypoint, link) (SURVEY A-1) ----------------------------------
        for (int li = tid; li < n_li; li += nthr) {
            const int i = li / NL, j = li - i * NL;
            double g[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) g[s] = 0.0;
            const float bv = s_best[li];
            if (j < jmax && bv > 0.0f && __float_as_uint(bv) >= tau) {
                const int p = s_bestp[li];
                const double *F = s_frames + (size_t)li * 12;
                const double *Fp = (i > 0) ? (F - NL * 12) : (s_frames + ((size_t)n * NL + j) * 12);
                const double *Fn = (i < n - 1) ? (F + NL * 12) : (s_frames + ((size_t)(n + 1) * NL + j) * 12);
                const double *bp = s_pts + ((size_t)j * P + p) * 3;
                double X, Y, Z, xp, yp, zp, xn, yn, zn;
                xform(F, bp[0], bp[1], bp[2], X, Y, Z);
                xform(Fp, bp[0], bp[1], bp[2], xp, yp, zp);
                xform(Fn, bp[0], bp[1], bp[2], xn, yn, zn);
                const float x = (float)X, y = (float)Y, z = (float)Z;
                float pot = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
                unsigned long long m = s_mask[li];
                while (m) {
                    const int o = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    float po, ax, ay, az, co;
                    pair_full(s_objs[o], a.grids, x, y, z, po, ax, ay, az, co);
                    pot = __fadd_rn(pot, po);
                    gx = __fadd_rn(gx, ax); gy = __fadd_rn(gy, ay); gz = __fadd_rn(gz, az);
                }
                if (finger_soft && j >= 8) {
                    pot = __fmul_rn(pot, 0.1f); gx = __fmul_rn(gx, 0.1f); gy = __fmul_rn(gy, 0.1f);
                    gz = __fmul_rn(gz, 0.1f);
                }
                functional_grad(s_jinfo + (size_t)i * NJ * 6, j, X, Y, Z, xp, yp, zp, xn, yn, zn, (double)pot,
                                (double)gx, (double)gy, (double)gz, dt, g);
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) s_lg[(size_t)li * NS + s] = g[s];
        }
    } else {
        obs_sum = block_sum(t_cost, s_red);
    }
    __syncthreads();

    // ---- phase 5: assemble gradient (cost.py:386-388/417-421, 425-449, 467-476) ----------------------
    const int jlast = (topk_mode && !prm.consider_finger) ? 7 : 9;
    double t_so = 0.0, t_ss = 0.0, t_sg = 0.0;
    for (int k = tid; k < n * ND; k += nthr) {
        const int i = k / ND, d = k - i * ND;
        double og = 0.0;
        if (d < 7) {
            for (int j = d; j <= jlast; ++j) og += s_lg[((size_t)i * NL + j) * NS + d];
        } else if (jlast == 9) {
            og = s_lg[((size_t)i * NL + (d + 1)) * NS + 7];   // dof 7 <- link 8, dof 8 <- link 9
        }
        const double xc = s_xi[k];
        const double xprev = (i > 0) ? s_xi[k - ND] : s_start[d];
        double sg;
        if (i < n - 1) sg = (2.0 * xc - xprev - s_xi[k + ND]) / (dt * dt);
        else sg = goal_set ? (xc - xprev) / (dt * dt) : (2.0 * xc - xprev - s_end[d]) / (dt * dt);
        sg *= prm.link_smooth_weight[d];
        double wo = prm.obstacle_weight * og;
        wo = fmin(fmax(wo, -prm.clip_grad_scale), prm.clip_grad_scale);
        const double ws = prm.smoothness_weight * sg;
        const double gt = wo + ws;
        s_grad[k] = gt;
        t_so += wo * wo; t_ss += ws * ws; t_sg += gt * gt;
    }
    // smoothness loss rows 0..n (cost.py:443-445)
    double t_sl = 0.0;
    for (int k = tid; k < (n + 1) * ND; k += nthr) {
        const int r = k / ND, d = k - r * ND;
        double v;
        if (r == 0) v = (s_xi[d] - s_start[d]) / dt;
        else if (r < n) v = (s_xi[k] - s_xi[k - ND]) / dt;
        else v = goal_set ? 0.0 : (s_end[d] - s_xi[(n - 1) * ND + d]) / dt;
        v *= prm.link_smooth_weight[d];
        t_sl += v * v;
    }
    const double norm_wo = sqrt(block_sum(t_so, s_red));
    const double norm_ws = sqrt(block_sum(t_ss, s_red));
    const double norm_g = sqrt(block_sum(t_sg, s_red));
    const double smooth_sum = 0.5 * block_sum(t_sl, s_red);
    double goal_dist = 0.0;
    if (goal_set) {
        double t = 0.0;
        if (tid < ND) { const double d = s_xi[(n - 1) * ND + tid] - s_end[tid]; t = d * d; }
        goal_dist = sqrt(block_sum(t, s_red));
    }
    // check_joint_limit (optimizer.py:166-174, sic: scalar "any below" times elementwise "above")
    double t_low = 0.0, t_high = 0.0;
    for (int k = tid; k < n * ND; k += nthr) {
        const int d = k % ND;
        if (s_xi[k] < rc->lower[d] - 5e-3) t_low = 1.0;
        if (s_xi[k] > rc->upper[d] + 5e-3) t_high = 1.0;
    }
    const bool any_low = block_sum(t_low, s_red) > 0.0;
    const bool any_high = block_sum(t_high, s_red) > 0.0;
    const bool violate = any_low && any_high;
    const bool terminate = (collide <= prm.allow_collision_point) && prm.pre_terminate && (goal_dist < 0.01) &&
                           (smooth_sum < prm.terminate_smooth_loss) && !violate;
    if (a.grad_out)
        for (int k = tid; k < n * ND; k += nthr) a.grad_out[(size_t)b * n * ND + k] = s_grad[k];

    int limit_rounds = 0;
    if (topk_mode && a.row_obs)
        for (int k = tid; k < n; k += nthr) a.row_obs[(size_t)b * n + k] = obs_sum / (double)n;
    // update: 0 = info only, 1 = always (force_update), 2 = unless this iteration reports terminate
    // (omg/optimizer.py:124-125)
    if (prm.update == 1 || (prm.update == 2 && !terminate)) {
        // ---- phase 6: covariant update -----------------------------------------------------------
        for (int k = tid; k < n * ND; k += nthr) {
            const int i = k / ND, d = k - i * ND;
            const double *Ar = a.Ainv + (size_t)i * n;
            double acc = 0.0;
            for (int r = 0; r < n; ++r) acc = fma(__ldg(Ar + r), s_grad[r * ND + d], acc);
            s_u[k] = acc;
        }
        __syncthreads();
        for (int k = tid; k < n * ND; k += nthr) {
            const int i = k / ND, d = k - i * ND;
            double up = -prm.step_size * s_u[k];
            if (goal_set) {
                double t1 = 0.0, t2 = 0.0;
                for (int r = 0; r < c; ++r) {
                    const double m = __ldg(a.proj + (size_t)i * c + r);
                    t1 = fma(m, s_u[(n - c + r) * ND + d], t1);
                    t2 = fma(m, s_xi[(n - c + r) * ND + d] - s_goal[r * ND + d], t2);
                }
                up = up + prm.step_size * t1 - t2;
            }
            s_viol[k] = up;
        }
        __syncthreads();
        for (int k = tid; k < n * ND; k += nthr) {   // Trajectory.update (core.py:43-51)
            const int d = k % ND;
            double v = s_xi[k];
            if (d < 7) v += s_viol[k];
            else v = fmin(fmax(v, 0.0), 0.04);
            s_xi[k] = v;
        }
        __syncthreads();
        // ---- phase 7: smooth joint-limit projection (optimizer.py:148-164) ------------------------
        for (int round = 0; round <= prm.joint_limit_max_steps; ++round) {
            double t = 0.0;
            for (int k = tid; k < n * ND; k += nthr) {
                const int d = k % ND;
                const double v = s_xi[k];
                double viol = 0.0;
                if (v < rc->lower[d]) viol = rc->lower[d] - v;
                else if (v > rc->upper[d]) viol = rc->upper[d] - v;
                s_viol[k] = viol;
                t += viol * viol;
            }
            const double vn = sqrt(block_sum(t, s_red));
            if (!(vn > 1e-2) || round == prm.joint_limit_max_steps) break;
            // argmax |viol|, first occurrence in flat order (np.argmax)
            double bm = -1.0;
            int bk = 0x7fffffff;
            for (int k = tid; k < n * ND; k += nthr) {
                const double m = fabs(s_viol[k]);
                if (m > bm) { bm = m; bk = k; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double om = __shfl_xor_sync(0xffffffffu, bm, off);
                const int ok = __shfl_xor_sync(0xffffffffu, bk, off);
                if (om > bm || (om == bm && ok < bk)) { bm = om; bk = ok; }
            }
            if (lane == 0) { s_red[warp] = bm; s_hist[warp] = bk; }
            __syncthreads();
            if (tid == 0) {
                double fm = s_red[0];
                int fk = s_hist[0];
                for (int w = 1; w < nwarps; ++w)
                    if (s_red[w] > fm || (s_red[w] == fm && s_hist[w] < fk)) { fm = s_red[w]; fk = s_hist[w]; }
                s_red[34] = fm;
                s_hist[260] = fk;
            }
            __syncthreads();
            const double vmax = s_red[34];
            const int kmax = s_hist[260];
            for (int k = tid; k < n * ND; k += nthr) {
                const int i = k / ND, d = k - i * ND;
                const double *Ar = a.Ainv + (size_t)i * n;
                double acc = 0.0;
                for (int r = 0; r < n; ++r) acc = fma(__ldg(Ar + r), s_viol[r * ND + d], acc);
                s_u[k] = acc;
            }
            __syncthreads();
            const double scale = vmax / (fabs(s_u[kmax]) + 1e-8);
            for (int k = tid; k < n * ND; k += nthr) s_xi[k] += scale * s_u[k];
            ++limit_rounds;
            __syncthreads();
        }
        // ---- phase 8: write back ---------------------------------------------------------------
        for (int k = tid; k < n * ND; k += nthr) g_xi[k] = s_xi[k];
    }
    if (tid == 0) {
        double *inf = a.info + (size_t)b * OMGB_INFO_STRIDE;
        inf[OMGB_INFO_OBS] = obs_sum;
        inf[OMGB_INFO_SMOOTH] = smooth_sum;
        inf[OMGB_INFO_COST] = prm.obstacle_weight * obs_sum + prm.smoothness_weight * smooth_sum;
        inf[OMGB_INFO_COLLIDE] = (double)collide;
        inf[OMGB_INFO_REACH] = goal_dist;
        inf[OMGB_INFO_GRAD_NORM] = norm_g;
        inf[OMGB_INFO_WOBS_GRAD_NORM] = norm_wo;
        inf[OMGB_INFO_WSMOOTH_GRAD_NORM] = norm_ws;
        inf[OMGB_INFO_TERMINATE] = terminate ? 1.0 : 0.0;
        inf[OMGB_INFO_VIOLATE_LIMIT] = violate ? 1.0 : 0.0;
        inf[OMGB_INFO_EXECUTE] =
            ((collide <= prm.allow_collision_point) && (smooth_sum < prm.terminate_smooth_loss)) ? 1.0 : 0.0;
        inf[OMGB_INFO_FAILURE_TERMINATE] = ((collide >= prm.allow_collision_point * 10) ||
                                            (smooth_sum >= prm.terminate_smooth_loss * 2.5)) ? 1.0 : 0.0;
        inf[OMGB_INFO_P_IN] = (double)p_in;
        inf[OMGB_INFO_NONZERO] = (double)nnz;
        inf[OMGB_INFO_LIMIT_ROUNDS] = (double)limit_rounds;
        inf[OMGB_INFO_RESERVED] = 0.0;
        if (a.done && a.stop_on_terminate && terminate && a.iteration > 0) a.done[b] = 1;
    }
}

}  // namespace omgb
