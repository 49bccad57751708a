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

#ifdef OMGB_NO_FP64_FMA
// Experiment only (tools/gpu_r02y.sh): every explicit fp64 multiply-add of the fused step rounded
// twice -- with -fmad=false for the implicit ones -- to see whether the 70-iteration outliers against the numpy oracle
// are an artefact of fused rounding.  The shipped library is built without this.
#define fma(a, b, c) __dadd_rn(__dmul_rn((a), (b)), (c))
#endif

namespace omgb {

constexpr int NL = OMGB_NUM_LINKS;
constexpr int ND = OMGB_NUM_DOF;
constexpr int NJ = 9;   // joint-info slots per waypoint: 7 arm joints, finger joint 8, finger joint 9
constexpr int NS = 8;   // gradient slots per link (<= 7 arm ancestors + own prismatic joint)

// Robot constants resident in HBM (read through L1; every access is warp-uniform or staged to smem).
struct RobotConst {
    double P0[10][12];     // pose_0[i]: rotation row-major [0:9], translation [9:12]
    double CO[10][12];     // center_offset[j]
    double jab[10][3];     // joint axis expressed in the link's BODY-POINT frame (T_j * center_offset_j)
    double job[10][3];     // joint "origin" (robot_pykdl.py:104 aliasing included) in the body-point frame
    double pts[10][OMGB_MAX_BODY_POINTS][3];
    float sph[10][4];      // bounding sphere of the link's body points (body-point frame centre, radius)
    double lower[ND], upper[ND];
    int p;                 // body points per link
    int pad_;
};

// Lower-bound grid over the packed SDFs: one float per 2x2x2-voxel brick = the minimum voxel value over the
// brick and its 26 neighbours (a 6^3-voxel region).  Any trilinear sample whose grid coordinate falls in the
// brick reads taps inside that region only, and a trilinear value is bounded below by its taps, so one load
// proves "this (body point, object) pair is farther than eps and not colliding" without touching the voxels.
struct DilDesc {
    const float *data;
    long long obj_stride;
    int bx, by, bz;
    int enabled;
};

struct SmemLayout {
    unsigned off_xi, off_start, off_end, off_goal, off_frames, off_lg, off_grad, off_u, off_viol, off_red,
        off_mask, off_mask_hi, off_best, off_bestp, off_act, off_win, off_objs, off_sph, off_hist, off_mbar, total;
    int nlu;        // links that own gradient rows: 8 in top-k mode without consider_finger (cost.py:401-402), else 10
    int red_max;    // first slot behind the sum buffers inside the reduction scratch
    int mask_hi;    // 1: more than 32 objects, the object masks are two 32-bit words
};

__host__ __device__ inline unsigned align_up(unsigned v, unsigned a) { return (v + a - 1) / a * a; }

// Shared memory of one CTA (one trajectory).  Sized so that a 30-waypoint trajectory fits three times and a
// 60-waypoint one twice into an SM's 228 KB (tests/host/step_layout_check.cu): the per-point potentials of the top-k path live in a global scratch
// (written and read once by the same CTA, L2-resident), link gradients only for the links that can own a winner,
// 32-bit object masks unless there are more than 32 objects, byte / short indices.
__host__ __device__ inline SmemLayout make_layout(int n, int c, int lpi, int nobj, int p, int nwarps, bool topk,
                                                  bool consider_finger) {
    SmemLayout L;
    unsigned o = 0;
    L.nlu = (topk && !consider_finger) ? NL - 2 : NL;
    L.mask_hi = nobj > 32 ? 1 : 0;
    L.off_xi = o; o += sizeof(double) * n * ND;
    L.off_start = o; o += sizeof(double) * ND;
    L.off_end = o; o += sizeof(double) * ND;
    L.off_goal = o; o += sizeof(double) * (c > 0 ? c : 1) * ND;
    L.off_frames = o; o += sizeof(double) * (n + 2) * NL * 12;
    // link gradients [n*nlu][8] fp64; aliased with the sin/cos table of the FK phase
    unsigned lg = sizeof(double) * n * L.nlu * NS, sc = sizeof(double2) * (n + 2) * 7;
    unsigned u = lg > sc ? lg : sc;
    o = align_up(o, 16);   // double2 sin/cos table
    L.off_lg = o; o += align_up(u, 16);
    // grad / u / viol (+ the scan scratch of metric_apply) live in the frames region: the link frames are dead once
    // the obstacle gradient is assembled; 5 * n * 9 doubles <= (n + 2) * 120
    L.off_grad = L.off_frames;
    L.off_u = L.off_grad + sizeof(double) * n * ND;
    L.off_viol = L.off_u + sizeof(double) * n * ND;
    // reduction scratch: two buffers of nwarps x 8 partial sums (block_sum_n), then from red_max on the per-warp and
    // final argmax slots of the joint-limit projection, later the 16-double info row
    L.red_max = 2 * nwarps * 8;
    L.off_red = o; o += sizeof(double) * (L.red_max + (nwarps + 1 < 16 ? 16 : nwarps + 1));
    L.off_mask = o; o += sizeof(unsigned) * n * NL;
    L.off_mask_hi = o; o += L.mask_hi ? sizeof(unsigned) * n * NL : 0;
    L.off_best = o; o += sizeof(float) * n * NL;
    L.off_act = o; o += align_up(sizeof(unsigned short) * (n * NL + 40), 4);
    L.off_win = o; o += align_up(sizeof(unsigned short) * (n * NL + 8), 4);
    L.off_bestp = o; o += align_up(n * NL, 16);
    o = align_up(o, 16);
    L.off_objs = o; o += sizeof(ObjRec) * nobj;
    L.off_sph = o; o += sizeof(float4) * nobj;
    L.off_hist = o; o += sizeof(int) * 264;
    L.off_mbar = o; o += 8;   // mbarrier of the bulk (TMA) staging copies
    L.total = align_up(o, 16);
    (void)p; (void)lpi;
    return L;
}

// Small robot constants passed as kernel parameters: they are read with warp-uniform indices inside dependent
// chains (FK, joint limits), where the constant bank avoids the L1/L2 round trip of a global load.
struct RobotParams {
    double P0[10][12];
    double CO[10][12];
    double lower[ND], upper[ND];
    float sph[10][4];
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
    const int *order;        // [B] or null: CTA -> trajectory map (longest-first scheduling hint; never changes results)
    int *cta_cost;           // [B] or null: clocks this trajectory's CTA took (feeds the next launch's order)
    long long *prof;         // [B,16] or null: clock64() at phase boundaries (diagnostic)
    double *hist_xi;         // [iters,B,n,9] or null: xi after every iteration (Planner.history_trajectories[1:])
    double *hist_info;       // [iters,B,16] or null: the info row of every iteration (Planner.info)
    QuadDesc quad;           // bricked quad copy of the grids for the exact path (data == null: reference layout)
    float *pot_scratch;      // [B,n*10,LPI] fp32: per-point potentials of the top-k path (global, written and read by
                             // the trajectory's own CTA within one iteration)
    DilDesc dil;
    SmemLayout lay;          // computed on the host (make_layout)
    RobotParams rp;
    int num_objects;
    int batch;
    int metric_kind;         // 0: dense Ainv; 1: Ainv = s*min(i,j) (goal-set, free end); 2: s*min(i,j)(n+1-max(i,j))/(n+1)
    double metric_scale;     // s (= dt^2)
    int bulk_stage;          // 1: xi and the object records are device memory -> staged by cp.async.bulk (TMA)
    int win_seg_min;         // phase 4b: from this many winners on (and more than blockDim / 4) pairs are evaluated one per lane
    int iteration;           // index inside a plan (for the t > 0 rule of planner.py:627)
    int stop_on_terminate;
    omgb_step_params_t prm;
};

// ----------------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------------
// TMA bulk copy global -> shared (cp.async.bulk, no tensor map: a contiguous run of 16-byte units) completing on an
// mbarrier: one thread moves a trajectory's xi rows and the scene's object records while the others stage the rest.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_one(unsigned bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses to the destination, too
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of K <= 8 values at once; every thread gets the results.  ONE barrier per call: the warp partials
// go to one of two scratch buffers of warps * 8 doubles, alternating from call to call (`flip`, kept in step by every
// thread of the CTA).  A partial written by call t + 2 cannot overtake a read of call t: its writer has passed the
// barrier of call t + 1, which every reader of call t reaches after its reads.  The partials are combined in warp
// order by lanes 0..K-1 of EVERY warp and broadcast by shuffles.
template <int K>
__device__ __forceinline__ void block_sum_n(double (&v)[K], double *scratch, int &flip) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double *buf = scratch + flip * nw * 8;
    flip ^= 1;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) buf[w * 8 + k] = v[k];
    }
    __syncthreads();
    double t = 0.0;
    if (lane < K)
        for (int q = 0; q < nw; ++q) t += buf[q * 8 + lane];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = __shfl_sync(0xffffffffu, t, k);
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

// Forward kinematics of one configuration q[9] (rad), one thread: 10 body-point frames
// (T_j * center_offset_j).  robot_pykdl.py:148-215; the rotX(+-pi)/column-flip pair of :166,174-176 cancels
// and is omitted.  Used by the batch-obstacle-cost kernel; the fused step uses the row-parallel form below.
__device__ void panda_fk(const RobotConst *__restrict__ rc, const double *q, double *frames) {
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
#pragma unroll
        for (int r = 0; r < 3; ++r) {   // ... * Rz(q_i): rotate the first two columns
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
    }
}

// One ROW of a rigid transform times a constant transform: the three rows of T_i = T_{i-1} * P0 * Rz evolve
// independently through the kinematic chain, so FK runs on 3 threads per configuration with no exchange.
struct Row { double a, b, c, t; };
__device__ __forceinline__ Row row_mul(const Row &r, const double *B) {
    Row o;
    o.a = fma(r.a, B[0], fma(r.b, B[3], r.c * B[6]));
    o.b = fma(r.a, B[1], fma(r.b, B[4], r.c * B[7]));
    o.c = fma(r.a, B[2], fma(r.b, B[5], r.c * B[8]));
    o.t = fma(r.a, B[9], fma(r.b, B[10], fma(r.c, B[11], r.t)));
    return o;
}

// frames: this configuration's [10][12]; row index r in 0..2; sc: sin/cos pairs of the 7 arm joints.
__device__ __forceinline__ void panda_fk_row(const RobotParams &rc, const double *q, const double2 *sc, int r,
                                             double *frames) {
    Row T;
    T.a = (r == 0) ? 1.0 : 0.0; T.b = (r == 1) ? 1.0 : 0.0; T.c = (r == 2) ? 1.0 : 0.0; T.t = 0.0;
    // (not unrolled: measured, the fully unrolled chain spills at the 64-register cap and the phase gets 40% slower)
#pragma unroll 1
    for (int i = 0; i < 7; ++i) {
        const Row N = row_mul(T, rc.P0[i]);
        const double s = sc[i].x, c = sc[i].y;
        T.a = fma(N.a, c, N.b * s);
        T.b = fma(N.b, c, -(N.a * s));
        T.c = N.c;
        T.t = N.t;
        const Row F = row_mul(T, rc.CO[i]);
        double *f = frames + 12 * i;
        f[3 * r] = F.a; f[3 * r + 1] = F.b; f[3 * r + 2] = F.c; f[9 + r] = F.t;
    }
    const Row H = row_mul(T, rc.P0[7]);
    {
        const Row F = row_mul(H, rc.CO[7]);
        double *f = frames + 12 * 7;
        f[3 * r] = F.a; f[3 * r + 1] = F.b; f[3 * r + 2] = F.c; f[9 + r] = F.t;
    }
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        const double *B = rc.P0[8 + k];
        const double ty = B[10] + ((k == 0) ? q[7] : -q[8]);   // robot_pykdl.py:181-184
        Row G;
        G.a = fma(H.a, B[0], fma(H.b, B[3], H.c * B[6]));
        G.b = fma(H.a, B[1], fma(H.b, B[4], H.c * B[7]));
        G.c = fma(H.a, B[2], fma(H.b, B[5], H.c * B[8]));
        G.t = fma(H.a, B[9], fma(H.b, ty, fma(H.c, B[11], H.t)));
        const Row F = row_mul(G, rc.CO[8 + k]);
        double *f = frames + 12 * (8 + k);
        f[3 * r] = F.a; f[3 * r + 1] = F.b; f[3 * r + 2] = F.c; f[9 + r] = F.t;
    }
}

// number of gradient slots of link j
__device__ __forceinline__ int link_slots(int j) { return j < 7 ? j + 1 : (j == 7 ? 7 : 8); }

// CHOMP functional gradient of one body point (omg/cost.py:24-43): the workspace vector w that the point
// Jacobian pulls back, w = |v| P grad_c - c P a / |v|^2 with P = I - v^ v^T.  x, xp, xn: the point at waypoint
// i, i-1, i+1.  Returns c * |v| (the point's obstacle cost).
__device__ __forceinline__ double fg_weight(double x, double y, double z, double xpx, double xpy, double xpz,
                                            double xnx, double xny, double xnz, double c, double gcx, double gcy,
                                            double gcz, double idt, double &wx, double &wy, double &wz) {
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
    wx = speed * (gcx - hx * hg) - ks * (ax - hx * ha);
    wy = speed * (gcy - hy * hg) - ks * (ay - hy * ha);
    wz = speed * (gcz - hz * hg) - ks * (az - hz * ha);
    return c * speed;
}

// Gradient slot s of link j: (point Jacobian column)^T w (omg/cost.py:92-110).  frames_i: the 10 body-point
// frames of waypoint i; joint axes and the reference's "origins" are rebuilt from them:
// axis_k = R_k * jab[k], origin_k = R_k * job[k] + t_k.
__device__ __forceinline__ double fg_slot(const RobotConst *__restrict__ rc, const double *frames_i, int j, int s,
                                          double x, double y, double z, double wx, double wy, double wz) {
    if (s >= link_slots(j)) return 0.0;
    const int k = (s < 7) ? s : j;            // joint id: arm joint s, or the finger's own joint
    const double *F = frames_i + 12 * k;
    const double a0 = rc->jab[k][0], a1 = rc->jab[k][1], a2 = rc->jab[k][2];
    const double ux = fma(F[0], a0, fma(F[1], a1, F[2] * a2));
    const double uy = fma(F[3], a0, fma(F[4], a1, F[5] * a2));
    const double uz = fma(F[6], a0, fma(F[7], a1, F[8] * a2));
    if (s < 7) {
        const double o0 = rc->job[k][0], o1 = rc->job[k][1], o2 = rc->job[k][2];
        const double rx = x - fma(F[0], o0, fma(F[1], o1, fma(F[2], o2, F[9])));
        const double ry = y - fma(F[3], o0, fma(F[4], o1, fma(F[5], o2, F[10])));
        const double rz = z - fma(F[6], o0, fma(F[7], o1, fma(F[8], o2, F[11])));
        return (uy * rz - uz * ry) * wx + (uz * rx - ux * rz) * wy + (ux * ry - uy * rx) * wz;
    }
    return ux * wx + uy * wy + uz * wz;   // prismatic finger joint: the column is the axis itself (cost.py:106-108)
}

__device__ __forceinline__ double functional_grad(const RobotConst *__restrict__ rc, const double *frames_i, int j,
                                                  double x, double y, double z, double xpx, double xpy, double xpz,
                                                  double xnx, double xny, double xnz, double c, double gcx,
                                                  double gcy, double gcz, double idt, double *g) {
    double wx, wy, wz;
    const double cost = fg_weight(x, y, z, xpx, xpy, xpz, xnx, xny, xnz, c, gcx, gcy, gcz, idt, wx, wy, wz);
#pragma unroll 1
    for (int s = 0; s < NS; ++s) g[s] = fg_slot(rc, frames_i, j, s, x, y, z, wx, wy, wz);
    return cost;
}

// Cheap classification of a (body point, object) pair from approximate (matrix-form, division-free) grid
// coordinates; their error is far below the margins used.
//   PAIR_FAR : the 8-tap cell is certainly in bounds AND the lower-bound grid proves value > eps and >= clearance
//              -> contributes nothing, counts in P_in;
//   PAIR_OUT : the 8-tap cell is certainly out of bounds -> the sample reads 1.0 (kernel.cu:47-48), contributes
//              nothing (eps < 1, clearance <= 1 is checked when the record is built), not in P_in;
//   PAIR_EXACT: anything else -> run the operator.
enum { PAIR_EXACT = 0, PAIR_FAR = 1, PAIR_OUT = 2, PAIR_LOAD = 3 };
// Stage 1: no memory access.  Returns PAIR_EXACT / PAIR_OUT, or PAIR_LOAD with the lower-bound entry to read.
__device__ __forceinline__ int classify_prepare(const ObjRec &ob, const DilDesc &dd, int oi, float x, float y,
                                                float z, const float *&addr) {
    (void)oi;
    const float gx = fmaf(ob.ga[0], x, fmaf(ob.ga[1], y, fmaf(ob.ga[2], z, ob.ga[3])));
    const float gy = fmaf(ob.ga[4], x, fmaf(ob.ga[5], y, fmaf(ob.ga[6], z, ob.ga[7])));
    const float gz = fmaf(ob.ga[8], x, fmaf(ob.ga[9], y, fmaf(ob.ga[10], z, ob.ga[11])));
    // in bounds <=> g in (-0.5, d - 0.5) on every axis (cell_of + the x1 < dim test of kernel.cu:47)
    const bool outside = (gx < -0.75f) | (gx > ob.fd0 - 0.25f) | (gy < -0.75f) | (gy > ob.fd1 - 0.25f) |
                         (gz < -0.75f) | (gz > ob.fd2 - 0.25f);
    const bool interior = (gx >= 1.5f) & (gx <= ob.fd0 - 1.5f) & (gy >= 1.5f) & (gy <= ob.fd1 - 1.5f) &
                          (gz >= 1.5f) & (gz <= ob.fd2 - 1.5f);
    const int ix = interior ? ((int)gx) >> 1 : 0, iy = interior ? ((int)gy) >> 1 : 0, iz = interior ? ((int)gz) >> 1 : 0;
    addr = dd.data + (ob.dil_off + (ix * dd.by + iy) * dd.bz + iz);   // 32-bit index (checked when the grid is built)
    if (outside) return (ob.cull_pad < 1e29f) ? PAIR_OUT : PAIR_EXACT;
    return interior ? PAIR_LOAD : PAIR_EXACT;
}
// Stage 2: the lower bound v proves "farther than eps and not colliding"?
__device__ __forceinline__ int classify_finish(const ObjRec &ob, float v) {
    const float slack = 1e-4f + 1e-5f * fabsf(v);   // an fp32 lerp may undershoot its taps by a few ulps
    return ((v > ob.eps + slack) & (v > ob.clr + slack)) ? PAIR_FAR : PAIR_EXACT;
}
__device__ __forceinline__ int classify_pair(const ObjRec &ob, const DilDesc &dd, int oi, float x, float y, float z) {
    const float *addr;
    const int st = classify_prepare(ob, dd, oi, x, y, z, addr);
    return st == PAIR_LOAD ? classify_finish(ob, __ldg(addr)) : st;
}

// First level of the cull, branch-free: (waypoint, link) sphere against every object's world-frame bounding sphere, one
// 128-bit shared-memory load per object from a compact copy of the spheres (sph_stage), independent iterations.  Most
// pairs end here; the survivors' bits go through the box tests.  Disabled objects sit at infinity; radius < 0: never cull.
__device__ __forceinline__ void sph_stage(const ObjRec *s_objs, float4 *s_sph, int O) {
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
        const ObjRec &ob = s_objs[o];
        const float inf = __int_as_float(0x7f800000);
        s_sph[o] = (ob.dis > 0.0f) ? make_float4(inf, inf, inf, 0.0f) : make_float4(ob.wsx, ob.wsy, ob.wsz, ob.wsr);
    }
}
__device__ __forceinline__ unsigned sph_near(const float4 *s_sph, int o0, int o1, float fx, float fy, float fz, float rad) {
    unsigned near = 0u;
#pragma unroll 4
    for (int o = o0; o < o1; ++o) {
        const float4 sp = s_sph[o];
        const float dx = fx - sp.x, dy = fy - sp.y, dz = fz - sp.z, rr = rad + sp.w;
        const bool far = (sp.w >= 0.0f) & (dx * dx + dy * dy + dz * dz > rr * rr);
        near |= (far ? 0u : 1u) << (o - o0);
    }
    return near;
}

// Phase 4b body: the winners of the top-k branch, G lanes per winner (the operator's 7 trilinear samples and the
// 8 gradient slots are spread over the group).
struct WinCtx {
    const RobotConst *rc;
    const double *frames;
    const unsigned *mask_lo, *mask_hi;   // (mask_hi null: at most 32 objects)
    const unsigned char *bestp;
    const unsigned short *win;
    const ObjRec *objs;
    const float *grids;
    const QuadDesc *quad;
    const DilDesc *dil;
    double *lg;
    double inv_dt;
    int n, n_win, nlu;
    bool finger_soft, use_dil;
};

template <int G>
__device__ __forceinline__ void winners_pass(const WinCtx &c) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    constexpr int GPWW = 32 / G;
    const int grp = lane / G, l = lane % G;
    const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
    for (int base = warp * GPWW; base < c.n_win; base += nwarps * GPWW) {
        const int idx = base + grp;
        if (idx >= c.n_win) continue;   // uniform within the group
        const int li = c.win[idx];
        const int i = li / NL, j = li - i * NL;
        const int p = c.bestp[li];
        const double *F = c.frames + (size_t)li * 12;
        const double *Fp = (i > 0) ? (F - NL * 12) : (c.frames + ((size_t)c.n * NL + j) * 12);
        const double *Fn = (i < c.n - 1) ? (F + NL * 12) : (c.frames + ((size_t)(c.n + 1) * NL + j) * 12);
        const double *bp = c.rc->pts[j][p];
        const double b0 = bp[0], b1 = bp[1], b2 = bp[2];
        double X, Y, Z, xp, yp, zp, xn, yn, zn;
        xform(F, b0, b1, b2, X, Y, Z);
        xform(Fp, b0, b1, b2, xp, yp, zp);
        xform(Fn, b0, b1, b2, xn, yn, zn);
        const float x = (float)X, y = (float)Y, z = (float)Z;
        float pot = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
#pragma unroll 1
        for (int half = 0; half < (c.mask_hi ? 2 : 1); ++half) {   // object mask: one or two 32-bit words
            unsigned m = half ? c.mask_hi[li] : c.mask_lo[li];
            while (m) {
                const int o = __ffs(m) - 1 + 32 * half;
                m &= m - 1;
                if (c.use_dil && classify_pair(c.objs[o], *c.dil, o, x, y, z) != PAIR_EXACT) continue;
                float po, ax, ay, az, co;
                pair_full_group<G>(c.objs[o], c.grids, *c.quad, gm, l, x, y, z, po, ax, ay, az, co);
                pot = __fadd_rn(pot, po);
                gx = __fadd_rn(gx, ax); gy = __fadd_rn(gy, ay); gz = __fadd_rn(gz, az);
            }
        }
        if (c.finger_soft && j >= 8) {
            pot = __fmul_rn(pot, 0.1f); gx = __fmul_rn(gx, 0.1f); gy = __fmul_rn(gy, 0.1f); gz = __fmul_rn(gz, 0.1f);
        }
        double wx, wy, wz;
        fg_weight(X, Y, Z, xp, yp, zp, xn, yn, zn, (double)pot, (double)gx, (double)gy, (double)gz, c.inv_dt, wx, wy, wz);
#pragma unroll 1
        for (int s = l; s < NS; s += G)
            c.lg[((size_t)i * c.nlu + j) * NS + s] = fg_slot(c.rc, c.frames + (size_t)i * NL * 12, j, s, X, Y, Z, wx, wy, wz);
    }
}

// Phase 4b for MANY winners: one lane per winner would walk its object mask serially and divergently (the heaviest
// trajectories: ~200 winners x up to 10 objects, the longest phase of the launch's critical path).  Here a warp takes a
// chunk of winners (lane = winner), classifies every (winner, object) pair with one lower-bound load, and then evaluates
// the surviving pairs ONE PER LANE, 32 at a time: pair t of the chunk belongs to the winner w with
// excl[w] <= t < incl[w] (prefix sums of the per-winner pair counts, found by a shuffle binary search) and is that
// winner's (t - excl[w])-th surviving object.  Results return to the owner by shuffles in pair order, so a winner's
// fp32 sums run over its objects in ascending order exactly as in the one-lane form.  No shared memory.
__device__ __forceinline__ void winners_pass_seg(const WinCtx &c) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int per = (c.n_win + nwarps - 1) / nwarps;   // winners per warp and chunk: all warps busy, at most one per lane
    per = per > 32 ? 32 : per;
    for (int cbase = warp * per; cbase < c.n_win; cbase += nwarps * per) {   // (warp-uniform)
        const int idx = cbase + lane;
        const bool have = lane < per && idx < c.n_win;
        const int li = have ? c.win[idx] : 0;
        const int i = li / NL, j = li - i * NL;
        const double *F = c.frames + (size_t)li * 12;
        const double *bp = c.rc->pts[j][have ? c.bestp[li] : 0];
        const double b0 = bp[0], b1 = bp[1], b2 = bp[2];
        double X, Y, Z;
        xform(F, b0, b1, b2, X, Y, Z);
        const float x = (float)X, y = (float)Y, z = (float)Z;
        unsigned nlo = 0u, nhi = 0u;
        if (have) {
            unsigned m = c.mask_lo[li];
            while (m) {
                const int o = __ffs(m) - 1;
                m &= m - 1;
                if (!c.use_dil || classify_pair(c.objs[o], *c.dil, o, x, y, z) == PAIR_EXACT) nlo |= 1u << o;
            }
            if (c.mask_hi) {
                m = c.mask_hi[li];
                while (m) {
                    const int o = __ffs(m) - 1;
                    m &= m - 1;
                    if (!c.use_dil || classify_pair(c.objs[o + 32], *c.dil, o + 32, x, y, z) == PAIR_EXACT) nhi |= 1u << o;
                }
            }
        }
        const int cnt = __popc(nlo) + __popc(nhi);
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const int excl = incl - cnt, total = __shfl_sync(0xffffffffu, incl, 31);
        float pot = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
        for (int t0 = 0; t0 < total; t0 += 32) {
            const bool on = t0 + lane < total;
            const int t = on ? t0 + lane : total - 1;
            int lo = 0, hi = 31;   // owner: the first lane whose inclusive count exceeds t
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int mid = (lo + hi) >> 1;
                if (__shfl_sync(0xffffffffu, incl, mid) > t) hi = mid;
                else lo = mid + 1;
            }
            const int w = lo;
            const int k = t - __shfl_sync(0xffffffffu, excl, w);
            const unsigned wl = __shfl_sync(0xffffffffu, nlo, w), wh = __shfl_sync(0xffffffffu, nhi, w);
            const float wx = __shfl_sync(0xffffffffu, x, w), wy = __shfl_sync(0xffffffffu, y, w),
                        wz = __shfl_sync(0xffffffffu, z, w);
            float po = 0.0f, ax = 0.0f, ay = 0.0f, az = 0.0f, co;
            if (on) {
                const int nl = __popc(wl);
                const int o = (k < nl) ? (int)__fns(wl, 0, k + 1) : 32 + (int)__fns(wh, 0, k - nl + 1);
                pair_full(c.objs[o], c.grids, *c.quad, wx, wy, wz, po, ax, ay, az, co);
            }
            // back to the owners, in pair order
            const int s0 = max(excl, t0) - t0, m = min(incl, t0 + 32) - t0 - s0;   // this batch holds m of my pairs from lane s0
            const int mmax = __reduce_max_sync(0xffffffffu, m);
            for (int r = 0; r < mmax; ++r) {
                const int src = (s0 + r) & 31;
                const float vp = __shfl_sync(0xffffffffu, po, src), vx = __shfl_sync(0xffffffffu, ax, src),
                            vy = __shfl_sync(0xffffffffu, ay, src), vz = __shfl_sync(0xffffffffu, az, src);
                if (r < m) {
                    pot = __fadd_rn(pot, vp);
                    gx = __fadd_rn(gx, vx); gy = __fadd_rn(gy, vy); gz = __fadd_rn(gz, vz);
                }
            }
        }
        if (c.finger_soft && j >= 8) {
            pot = __fmul_rn(pot, 0.1f); gx = __fmul_rn(gx, 0.1f); gy = __fmul_rn(gy, 0.1f); gz = __fmul_rn(gz, 0.1f);
        }
        if (have) {
            const double *Fp = (i > 0) ? (F - NL * 12) : (c.frames + ((size_t)c.n * NL + j) * 12);
            const double *Fn = (i < c.n - 1) ? (F + NL * 12) : (c.frames + ((size_t)(c.n + 1) * NL + j) * 12);
            double xp, yp, zp, xn, yn, zn, wx, wy, wz;
            xform(Fp, b0, b1, b2, xp, yp, zp);
            xform(Fn, b0, b1, b2, xn, yn, zn);
            fg_weight(X, Y, Z, xp, yp, zp, xn, yn, zn, (double)pot, (double)gx, (double)gy, (double)gz, c.inv_dt, wx, wy, wz);
#pragma unroll 1
            for (int s = 0; s < NS; ++s)
                c.lg[((size_t)i * c.nlu + j) * NS + s] = fg_slot(c.rc, c.frames + (size_t)i * NL * 12, j, s, X, Y, Z, wx, wy, wz);
        }
    }
}

// dst = Ainv * src for [n][9] arrays in shared memory (all threads call; ends with a barrier).  The smoothness metric
// A = K^T K is tridiagonal, so its inverse has the closed forms above and Ainv * g is two running
// sums per DOF column instead of a dense n x n product: 9 threads scan, everyone combines.  tmp: [2][n][9] scratch.
__device__ __forceinline__ void metric_apply(const StepArgs &a, int n, const double *src, double *dst, double *tmp) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (a.metric_kind == 0) {
        for (int k = tid; k < n * ND; k += nthr) {
            const int i = k / ND, d = k - i * ND;
            const double *Ar = a.Ainv + (size_t)i * n;
            double acc = 0.0;
            for (int r = 0; r < n; ++r) acc = fma(__ldg(Ar + r), src[r * ND + d], acc);
            dst[k] = acc;
        }
        __syncthreads();
        return;
    }
    double *P = tmp, *S = tmp + n * ND;   // P_i = sum_{j<=i} j g_j, S_i = sum_{j<=i} g_j  (1-based i, j)
    {   // one warp per DOF column: shuffle scans over the waypoints, 32 at a time with a carry
        const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
        for (int d = warp; d < ND; d += nwarps) {
            double carry_p = 0.0, carry_s = 0.0;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                const double g = (i < n) ? src[i * ND + d] : 0.0;
                double p = (double)(i + 1) * g, sa = g;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const double tp = __shfl_up_sync(0xffffffffu, p, off), ts = __shfl_up_sync(0xffffffffu, sa, off);
                    if (lane >= off) { p += tp; sa += ts; }
                }
                p += carry_p; sa += carry_s;
                if (i < n) { P[i * ND + d] = p; S[i * ND + d] = sa; }
                carry_p = __shfl_sync(0xffffffffu, p, 31); carry_s = __shfl_sync(0xffffffffu, sa, 31);
            }
        }
    }
    __syncthreads();
    const double np1 = (double)(n + 1);
    for (int k = tid; k < n * ND; k += nthr) {
        const int i = k / ND, d = k - i * ND;
        const double ii = (double)(i + 1);
        const double Pn = P[(n - 1) * ND + d], Sn = S[(n - 1) * ND + d];
        const double tailS = Sn - S[k];
        double v;
        if (a.metric_kind == 1) v = P[k] + ii * tailS;
        else v = ((np1 - ii) * P[k] + ii * (np1 * tailS - (Pn - P[k]))) / np1;
        dst[k] = a.metric_scale * v;
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------------------------------
// the fused iteration
// ----------------------------------------------------------------------------------------------------
// One CHOMP iteration of trajectory b by the calling CTA (every thread calls).  w_obs / w_smooth / step_size: the
// schedules Optimizer.update() set for this iteration (omg/optimizer.py:59-80); iteration: index inside a plan.
template <int LPI, bool TOPK>
__device__ __forceinline__ void chomp_iteration(const StepArgs &a, unsigned char *smem, const int b, const double w_obs,
                                                const double w_smooth, const double step_size, const int iteration,
                                                unsigned &bulk_uses) {
    const long long t_begin = clock64();

    const omgb_step_params_t &prm = a.prm;
    const int n = prm.n_waypoints, c = prm.constraint_rows;
    const RobotConst *__restrict__ rc = a.robot;
    const int P = rc->p;
    const int O = a.num_objects;
    const SmemLayout &L = a.lay;
    double *s_xi = reinterpret_cast<double *>(smem + L.off_xi);
    double *s_start = reinterpret_cast<double *>(smem + L.off_start);
    double *s_end = reinterpret_cast<double *>(smem + L.off_end);
    double *s_goal = reinterpret_cast<double *>(smem + L.off_goal);
    double *s_frames = reinterpret_cast<double *>(smem + L.off_frames);
    double *s_lg = reinterpret_cast<double *>(smem + L.off_lg);
    double2 *s_sc = reinterpret_cast<double2 *>(smem + L.off_lg);
    double *s_grad = reinterpret_cast<double *>(smem + L.off_grad);
    double *s_u = reinterpret_cast<double *>(smem + L.off_u);
    double *s_viol = reinterpret_cast<double *>(smem + L.off_viol);
    double *s_red = reinterpret_cast<double *>(smem + L.off_red);
    unsigned *s_mlo = reinterpret_cast<unsigned *>(smem + L.off_mask);
    unsigned *s_mhi = L.mask_hi ? reinterpret_cast<unsigned *>(smem + L.off_mask_hi) : nullptr;
    float *s_best = reinterpret_cast<float *>(smem + L.off_best);
    unsigned char *s_bestp = smem + L.off_bestp;
    unsigned short *s_act = reinterpret_cast<unsigned short *>(smem + L.off_act);
    unsigned short *s_win = reinterpret_cast<unsigned short *>(smem + L.off_win);
    ObjRec *s_objs = reinterpret_cast<ObjRec *>(smem + L.off_objs);
    float4 *s_sph = reinterpret_cast<float4 *>(smem + L.off_sph);
    int *s_hist = reinterpret_cast<int *>(smem + L.off_hist);

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    int bs_flip = 0;   // block_sum_n's scratch buffer of the next call
    const double dt = prm.time_interval;
    const double inv_dt = 1.0 / dt, inv_dt2 = inv_dt * inv_dt;   // fp64 divisions are ~50 instructions each
    constexpr bool topk_mode = TOPK;   // prm.top_k_collision > 0
    const bool goal_set = prm.goal_set_proj != 0;
    const int n_li = n * NL;
    const int NLU = L.nlu;
    float *g_pot = a.pot_scratch + (size_t)b * n_li * LPI;   // this trajectory's slice of the potential scratch

#define OMGB_PROF(slot) do { if (a.prof && tid == 0) a.prof[(size_t)b * 16 + (slot)] = clock64(); } while (0)
    OMGB_PROF(0);
    if (a.prof && tid == 0) {   // diagnostic timeline: SM id and global-timer start / end of this CTA
        unsigned smid;
        unsigned long long gt;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.prof[(size_t)b * 16 + 13] = (long long)smid;
        a.prof[(size_t)b * 16 + 14] = (long long)gt;
    }
    // ---- phase 0: stage ---------------------------------------------------------------------------
    // xi, start, end and the goal rows may live in mapped pinned HOST memory (omgb_chomp_step_host, zero-copy): every
    // load of them is a PCIe round trip, so all of a thread's loads are issued back to back into registers before
    // the first dependent store (s_xi, s_start, s_end, s_goal are contiguous in shared memory, in this order), and
    // the device-memory staging below runs while they are in flight.
    double *g_xi = a.xi + (size_t)b * n * ND;
    const int n_x = n * ND, n_in = n_x + 2 * ND + c * ND;
    // (.cg loads: in the persistent plan kernel the previous iteration of this trajectory may have run on another SM)
    auto stage_src = [&](int k) -> const double * {
        if (k < n_x) return g_xi + k;
        if (k < n_x + ND) return a.start + (size_t)b * ND + (k - n_x);
        if (k < n_x + 2 * ND) return a.end + (size_t)b * ND + (k - n_x - ND);
        return a.goal_rows + (size_t)b * c * ND + (k - n_x - 2 * ND);
    };
    // Device-resident state: the xi rows (n*72 bytes) and the object records (288 bytes each) are contiguous runs of
    // 16-byte units -> ONE thread issues two TMA bulk copies that complete on an mbarrier (they read through L2, which
    // is where another SM's previous iteration of this trajectory left xi); start / end / goal rows (72-byte rows,
    // not 16-byte aligned for odd b) and everything in mapped host memory take the per-thread path.
    const unsigned mbar = smem_u32(smem + L.off_mbar);
    const bool bulk = a.bulk_stage && ((((size_t)g_xi) | ((size_t)n_x * 8u)) & 15u) == 0;
    if (bulk && tid == 0) {
        const unsigned bytes_xi = (unsigned)n_x * 8u, bytes_obj = (unsigned)sizeof(ObjRec) * (unsigned)O;
        // the mbarrier is initialised once per CTA; a persistent CTA's later items complete its next phases (the
        // generic-proxy accesses of the previous item to the destinations are ordered before the copies by the fence)
        if (bulk_uses == 0u) mbar_init_one(mbar);
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(mbar, bytes_xi + bytes_obj);
        bulk_g2s(smem_u32(s_xi), g_xi, bytes_xi, mbar);
        bulk_g2s(smem_u32(s_objs), a.objs, bytes_obj, mbar);
    }
    const int k0 = bulk ? n_x : 0;
    double stg0 = 0.0, stg1 = 0.0;
    if (k0 + tid < n_in) stg0 = __ldcg(stage_src(k0 + tid));
    if (k0 + tid + nthr < n_in) stg1 = __ldcg(stage_src(k0 + tid + nthr));
    if (!bulk) {
        const int words = (int)(sizeof(ObjRec) / 4) * O;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.objs);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_objs);
        for (int k = tid; k < words; k += nthr) dst[k] = src[k];
    }
    for (int k = tid; k < n_li; k += nthr) { s_best[k] = 0.0f; s_bestp[k] = 0; }
    if (k0 + tid < n_in) s_xi[k0 + tid] = stg0;
    if (k0 + tid + nthr < n_in) s_xi[k0 + tid + nthr] = stg1;
    for (int k = k0 + tid + 2 * nthr; k < n_in; k += nthr) s_xi[k] = __ldcg(stage_src(k));   // (long trajectories)
    __syncthreads();
    if (bulk) { mbar_wait(mbar, bulk_uses & 1u); ++bulk_uses; }
    sph_stage(s_objs, s_sph, O);   // (read after the barriers of the FK phase)

    OMGB_PROF(1);
    // ---- phase 1: forward kinematics (n waypoints, then start, then end) ---------------------------
    // 1a: sin/cos of every arm joint angle, one per thread
    for (int k = tid; k < (n + 2) * 7; k += nthr) {
        const int cfg = k / 7, i = k - cfg * 7;
        const double *q = (cfg < n) ? (s_xi + cfg * ND) : (cfg == n ? s_start : s_end);
        double sn, cs;
        sincos(q[i], &sn, &cs);
        s_sc[k] = make_double2(sn, cs);
    }
    __syncthreads();
    // 1b: three threads per configuration, one transform row each
    for (int k = tid; k < (n + 2) * 3; k += nthr) {
        const int cfg = k / 3, r = k - cfg * 3;
        const double *q = (cfg < n) ? (s_xi + cfg * ND) : (cfg == n ? s_start : s_end);
        panda_fk_row(a.rp, q, s_sc + cfg * 7, r, s_frames + (size_t)cfg * NL * 12);
    }
    __syncthreads();

    OMGB_PROF(2);
    // ---- phase 1c: cull.  (waypoint, link) bounding sphere vs each object:
    //  (i)  outside the region of the grid in which an 8-tap cell is in bounds: every sample reads 1.0 and
    //       contributes nothing (kernel.cu:47-48);
    //  (ii) strictly inside the grid but outside the object's ACTIVE box (the AABB of all bricks whose lower
    //       bound allows value <= eps or < clearance): nothing contributes, all P pairs count as in-bounds. ----
    int t_pin = 0;
    const bool use_dil = a.dil.enabled != 0;
    for (int li = tid; li < n_li; li += nthr) {
        const int j = li % NL;
        double cx, cy, cz;
        xform(s_frames + (size_t)li * 12, (double)a.rp.sph[j][0], (double)a.rp.sph[j][1], (double)a.rp.sph[j][2], cx, cy, cz);
        const float fx = (float)cx, fy = (float)cy, fz = (float)cz, rad = a.rp.sph[j][3];
        unsigned long long m = 0ull;
        // first level: link sphere vs the objects' world-frame bounding spheres
        unsigned long long near = sph_near(s_sph, 0, O < 32 ? O : 32, fx, fy, fz, rad);
        if (O > 32) near |= (unsigned long long)sph_near(s_sph, 32, O, fx, fy, fz, rad) << 32;
        while (near) {
            const int o = __ffsll((long long)near) - 1;
            near &= near - 1;
            const ObjRec &ob = s_objs[o];
            if (ob.dis > 0.0f) continue;
            const float qx = ob.r[0] * fx + ob.r[1] * fy + ob.r[2] * fz + ob.tx;
            const float qy = ob.r[3] * fx + ob.r[4] * fy + ob.r[5] * fz + ob.ty;
            const float qz = ob.r[6] * fx + ob.r[7] * fy + ob.r[8] * fz + ob.tz;
            const float s = rad + ob.cull_pad;
            const bool box = (qx > ob.lox - s) & (qx < ob.hix + s) & (qy > ob.loy - s) & (qy < ob.hiy + s) &
                             (qz > ob.loz - s) & (qz < ob.hiz + s);
            if (!box) continue;
            if (use_dil && ob.cull_pad < 1e29f && !a.dbg_pts) {
                const bool miss = (qx + s < ob.alox) | (qx - s > ob.ahix) | (qy + s < ob.aloy) | (qy - s > ob.ahiy) |
                                  (qz + s < ob.aloz) | (qz - s > ob.ahiz);
                if (miss) {
                    const bool interior =
                        ((qx - s - ob.minx) * ob.isx >= 1.5f) & ((qx + s - ob.minx) * ob.isx <= ob.fd0 - 1.5f) &
                        ((qy - s - ob.miny) * ob.isy >= 1.5f) & ((qy + s - ob.miny) * ob.isy <= ob.fd1 - 1.5f) &
                        ((qz - s - ob.minz) * ob.isz >= 1.5f) & ((qz + s - ob.minz) * ob.isz <= ob.fd2 - 1.5f);
                    if (interior) { t_pin += P; continue; }
                }
            }
            m |= 1ull << o;
        }
        s_mlo[li] = (unsigned)m;
        if (s_mhi) s_mhi[li] = (unsigned)(m >> 32);
    }
    __syncthreads();
    OMGB_PROF(3);
    // ordered compaction of the link instances that still have work
    {
        int n_act = 0;   // running count (uniform)
        for (int base = 0; base < n_li; base += nthr) {
            const int li = base + tid;
            const bool on = (li < n_li) && (s_mlo[li] != 0u || (s_mhi && s_mhi[li] != 0u) || a.dbg_pts != nullptr);
            const unsigned bal = __ballot_sync(0xffffffffu, on);
            if (lane == 0) s_hist[warp] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < nwarps; ++w) {
                const int cnt = s_hist[w];
                if (w < warp) before += cnt;
                total += cnt;
            }
            if (on) s_act[n_act + before + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)li;
            n_act += total;
            __syncthreads();
        }
        if (tid == 0) s_hist[263] = n_act;
    }
    if (!topk_mode)
        for (int k = tid; k < n_li * NS; k += nthr) s_lg[k] = 0.0;   // (the sin/cos table is dead now)
    __syncthreads();
    const int n_act = s_hist[263];

    OMGB_PROF(4);
    // ---- phase 2: body points x objects ------------------------------------------------------------
    constexpr int GPW = 32 / LPI;                 // link instances per warp
    const int sub = lane / LPI, pl = lane % LPI;  // which instance of the warp, which body point
    const unsigned gmask = (LPI == 32) ? 0xffffffffu : (0xffffu << (sub * 16));
    const bool finger_soft = (prm.uncheck_finger_collision == -1);
    const int jmax = (topk_mode && !prm.consider_finger) ? NL - 2 : NL;   // links that count towards cost / gradient
    int t_nnz = 0, t_col = 0, t_exact = 0;
    double t_cost = 0.0;
    for (int base = warp * GPW; base < n_act; base += nwarps * GPW) {
        const int idx = base + sub;
        const bool have = idx < n_act;
        const int li = s_act[have ? idx : n_act - 1];
        const bool live = have && (pl < P);
        const int i = li / NL, j = li - i * NL;
        const double *F = s_frames + (size_t)li * 12;
        const double *bp = rc->pts[j][pl < P ? pl : 0];
        double X, Y, Z;
        xform(F, bp[0], bp[1], bp[2], X, Y, Z);
        const float x = (float)X, y = (float)Y, z = (float)Z;   // omg/cost.py:136 .float()
        float pot = 0.0f, col = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
#pragma unroll 1
        for (int half = 0; half < (s_mhi ? 2 : 1); ++half) {   // object mask: one or two 32-bit words
            unsigned m = live ? (half ? s_mhi[li] : s_mlo[li]) : 0u;
            while (m) {
                const int o = __ffs(m) - 1 + 32 * half;
                m &= m - 1;
                float po, co;
                bool inb;
                if (use_dil) {
                    const int cls = classify_pair(s_objs[o], a.dil, o, x, y, z);
                    if (cls != PAIR_EXACT) {   // provably contributes nothing
                        t_pin += (cls == PAIR_FAR) ? 1 : 0;
                        continue;
                    }
                }
                t_exact += 1;
                if (topk_mode) {
                    inb = pair_potential(s_objs[o], a.grids, a.quad, x, y, z, po, co);
                } else {
                    float ax, ay, az;
                    inb = pair_full(s_objs[o], a.grids, a.quad, x, y, z, po, ax, ay, az, co);
                    gx = __fadd_rn(gx, ax); gy = __fadd_rn(gy, ay); gz = __fadd_rn(gz, az);
                }
                pot = __fadd_rn(pot, po);
                col = __fadd_rn(col, co);
                t_pin += inb ? 1 : 0;
            }
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
            if (live) __stcg(g_pot + (size_t)li * LPI + pl, pot);
            // argmax over the body points of this link instance (potentials are >= 0, so their bit patterns
            // order like the values); ties -> highest point index
            const unsigned bits = live ? __float_as_uint(pot) : 0u;
            const unsigned mx = __reduce_max_sync(gmask, bits);
            const unsigned bal = __ballot_sync(gmask, live && bits == mx) & gmask;
            if (have && pl == 0 && mx != 0u) {
                s_best[li] = __uint_as_float(mx);
                s_bestp[li] = (unsigned char)((31 - __clz(bal)) - sub * LPI);
            }
            // c * |v| of every non-zero point: it IS the obstacle cost when at most k points are non-zero (the usual
            // case: every non-zero point is a member, cost.py:390-397); otherwise phase 4a sums the members
            if (live && pot > 0.0f && j < jmax) {
                const double *Fp = (i > 0) ? (F - NL * 12) : (s_frames + ((size_t)n * NL + j) * 12);
                double xp, yp, zp;
                xform(Fp, bp[0], bp[1], bp[2], xp, yp, zp);
                const double vx = (X - xp) * inv_dt, vy = (Y - yp) * inv_dt, vz = (Z - zp) * inv_dt;
                t_cost += (double)pot * sqrt(vx * vx + vy * vy + vz * vz);
            }
        } else {
            // full-sum mode: functional gradient of every point with non-zero potential, reduced over
            // the link instance's body points with warp shuffles
            double g[NS];
            double cst = 0.0;
#pragma unroll
            for (int s = 0; s < NS; ++s) g[s] = 0.0;
            const bool nz = live && (pot != 0.0f || gx != 0.0f || gy != 0.0f || gz != 0.0f);
            if (nz) {
                const double *Fp = (i > 0) ? (F - NL * 12) : (s_frames + ((size_t)n * NL + j) * 12);
                const double *Fn = (i < n - 1) ? (F + NL * 12) : (s_frames + ((size_t)(n + 1) * NL + j) * 12);
                double xp, yp, zp, xn, yn, zn;
                xform(Fp, bp[0], bp[1], bp[2], xp, yp, zp);
                xform(Fn, bp[0], bp[1], bp[2], xn, yn, zn);
                cst = functional_grad(rc, s_frames + (size_t)i * NL * 12, j, X, Y, Z, xp, yp, zp, xn, yn, zn,
                                      (double)pot, (double)gx, (double)gy, (double)gz, inv_dt, g);
            }
            if (__ballot_sync(gmask, nz) & gmask) {   // uniform per link instance
#pragma unroll
                for (int off = LPI / 2; off > 0; off >>= 1) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) g[s] += __shfl_xor_sync(gmask, g[s], off, LPI);
                    cst += __shfl_xor_sync(gmask, cst, off, LPI);
                }
                if (have && pl == 0) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) s_lg[(size_t)li * NS + s] = g[s];
                    t_cost += cst;
                    if (a.row_obs) atomicAdd(a.row_obs + (size_t)b * n + i, cst);
                }
            }
        }
    }
    OMGB_PROF(5);
    double red4[4] = {(double)t_nnz, (double)t_pin, (double)t_col, t_cost};
    block_sum_n<4>(red4, s_red, bs_flip);
    if (a.prof) {   // diagnostic: exact operator evaluations in the points phase
        double ex[1] = {(double)t_exact};
        block_sum_n<1>(ex, s_red, bs_flip);
        if (tid == 0) a.prof[(size_t)b * 16 + 12] = (long long)(ex[0] + 0.5);
    }
    const int nnz = (int)(red4[0] + 0.5), p_in = (int)(red4[1] + 0.5), collide = (int)(red4[2] + 0.5);

    double obs_sum = 0.0;
    if (topk_mode) {
        // ---- phase 3: membership threshold = k-th largest potential (bit pattern; potentials >= 0) ----
        const int K = prm.top_k_collision;
        uint32_t tau = 1u;   // "pot >= tau" <=> pot > 0
        double acc = 0.0;
        if (nnz > K) {
            uint32_t prefix = 0u, pmask = 0u;
            int remaining = K;
            // each thread keeps its share of the active slots' bit patterns in registers across the four passes
            constexpr int RC = 16;
            uint32_t cache[RC];
            const int n_as = n_act * LPI;
#pragma unroll
            for (int q = 0; q < RC; ++q) {
                const int k = tid + q * nthr;
                cache[q] = (k < n_as && (k % LPI) < P) ? __float_as_uint(__ldcg(g_pot + (size_t)s_act[k / LPI] * LPI + (k % LPI))) : 0u;
            }
            for (int shift = 24; shift >= 0; shift -= 8) {
                for (int k = tid; k < 256; k += nthr) s_hist[k] = 0;
                __syncthreads();
#pragma unroll
                for (int q = 0; q < RC; ++q) {
                    const uint32_t u = cache[q];
                    if (u != 0u && (u & pmask) == prefix) atomicAdd(&s_hist[(u >> shift) & 255u], 1);
                }
                for (int k = tid + RC * nthr; k < n_as; k += nthr) {   // (only when n_as > 16 * blockDim)
                    if ((k % LPI) >= P) continue;
                    const uint32_t u = __float_as_uint(__ldcg(g_pot + (size_t)s_act[k / LPI] * LPI + (k % LPI)));
                    if (u != 0u && (u & pmask) == prefix) atomicAdd(&s_hist[(u >> shift) & 255u], 1);
                }
                __syncthreads();
                if (warp == 0) {   // descending scan of the 256 buckets: 8 per lane
                    int loc[8], sum = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) { loc[q] = s_hist[255 - (lane * 8 + q)]; sum += loc[q]; }
                    int incl = sum;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, off);
                        if (lane >= off) incl += t;
                    }
                    const int excl = incl - sum;
                    if (excl < remaining && incl >= remaining) {
                        int acc = excl;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (acc + loc[q] >= remaining) {
                                s_hist[256] = 255 - (lane * 8 + q);
                                s_hist[257] = remaining - acc;
                                break;
                            }
                            acc += loc[q];
                        }
                    }
                }
                __syncthreads();
                prefix |= ((uint32_t)s_hist[256]) << shift;
                pmask |= 255u << shift;
                remaining = s_hist[257];
                __syncthreads();
            }
            tau = prefix;   // ties at tau are all kept (reference: unstable argsort, order undefined)
            // ---- phase 4a: obstacle cost = sum over members of c*|v| (cost.py:30,416), links 0..7; only when more than
            // k points are non-zero (else the points phase has summed it already).  The potentials are still in the
            // registers the selection kept them in; same slots per thread, same order as a sweep of the scratch ----
            auto member_cost = [&](int k, uint32_t u) {
                if (u == 0u || u < tau) return;
                const int li = s_act[k / LPI], p = k % LPI;
                const int i = li / NL, j = li - i * NL;
                if (j >= jmax) return;
                const double *F = s_frames + (size_t)li * 12;
                const double *Fp = (i > 0) ? (F - NL * 12) : (s_frames + ((size_t)n * NL + j) * 12);
                const double *bp = rc->pts[j][p];
                double X, Y, Z, xp, yp, zp;
                xform(F, bp[0], bp[1], bp[2], X, Y, Z);
                xform(Fp, bp[0], bp[1], bp[2], xp, yp, zp);
                const double vx = (X - xp) * inv_dt, vy = (Y - yp) * inv_dt, vz = (Z - zp) * inv_dt;
                acc += (double)__uint_as_float(u) * sqrt(vx * vx + vy * vy + vz * vz);
            };
#pragma unroll
            for (int q = 0; q < RC; ++q) member_cost(tid + q * nthr, cache[q]);
            for (int k = tid + RC * nthr; k < n_as; k += nthr) {   // (only when n_as > 16 * blockDim)
                if ((k % LPI) >= P) continue;
                member_cost(k, __float_as_uint(__ldcg(g_pot + (size_t)s_act[k / LPI] * LPI + (k % LPI))));
            }
        }
        OMGB_PROF(6);
        // winner list: active link instances whose best point is a top-k member (unordered; each winner writes
        // only its own gradient rows)
        if (tid == 0) s_hist[261] = 0;
        __syncthreads();
        for (int base = 0; base < n_act; base += nthr) {
            const int idx = base + tid;
            bool win = false;
            int li = 0;
            if (idx < n_act) {
                li = s_act[idx];
                const float bv = s_best[li];
                win = ((li % NL) < jmax) && (bv > 0.0f) && (__float_as_uint(bv) >= tau);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, win);
            int wbase = 0;
            if (lane == 0 && bal) wbase = atomicAdd(&s_hist[261], __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (win) s_win[wbase + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)li;
        }
        obs_sum = red4[3];
        if (nnz > K) {   // (uniform)
            double red1[1] = {acc};
            block_sum_n<1>(red1, s_red, bs_flip);
            obs_sum = red1[0];
        }
        obs_sum *= (double)n;   // added to every waypoint row (SURVEY A-3)
        for (int k = tid; k < n * NLU * NS; k += nthr) s_lg[k] = 0.0;
        __syncthreads();
        OMGB_PROF(7);
        // ---- phase 4b: one winner per (waypoint, link) (SURVEY A-1).  The winner list was compacted in phase 4a;
        // the lanes per winner adapt to how many there are (few winners: low latency; many: no redundancy) ----
        {
            WinCtx wc;
            wc.rc = rc; wc.frames = s_frames; wc.mask_lo = s_mlo; wc.mask_hi = s_mhi; wc.bestp = s_bestp; wc.win = s_win;
            wc.objs = s_objs; wc.nlu = NLU;
            wc.grids = a.grids; wc.quad = &a.quad; wc.dil = &a.dil; wc.lg = s_lg; wc.inv_dt = inv_dt; wc.n = n; wc.n_win = s_hist[261];
            wc.finger_soft = finger_soft; wc.use_dil = use_dil;
            if (wc.n_win * 8 <= nthr) winners_pass<8>(wc);
            else if (wc.n_win * 4 <= nthr) winners_pass<4>(wc);
            else if (wc.n_win < a.win_seg_min) {
                if (wc.n_win * 2 <= nthr) winners_pass<2>(wc);
                else winners_pass<1>(wc);
            } else winners_pass_seg(wc);
        }
    } else {
        obs_sum = red4[3];
    }
    __syncthreads();

    OMGB_PROF(8);
    // ---- phase 5: assemble gradient (cost.py:386-388/417-421, 425-449, 467-476) ----------------------
    const int jlast = (topk_mode && !prm.consider_finger) ? 7 : 9;
    double red7[7] = {0, 0, 0, 0, 0, 0, 0};   // |w_obs g|^2, |w_smooth g|^2, |g|^2, smooth rows, goal^2, low, high
    for (int k = tid; k < n * ND; k += nthr) {
        const int i = k / ND, d = k - i * ND;
        double og = 0.0;
        if (d < 7) {
            for (int j = d; j <= jlast; ++j) og += s_lg[((size_t)i * NLU + j) * NS + d];
        } else if (jlast == 9) {
            og = s_lg[((size_t)i * NLU + (d + 1)) * NS + 7];   // dof 7 <- link 8, dof 8 <- link 9
        }
        const double xc = s_xi[k];
        const double xprev = (i > 0) ? s_xi[k - ND] : s_start[d];
        double sg;
        if (i < n - 1) sg = (2.0 * xc - xprev - s_xi[k + ND]) * inv_dt2;
        else sg = goal_set ? (xc - xprev) * inv_dt2 : (2.0 * xc - xprev - s_end[d]) * inv_dt2;
        sg *= prm.link_smooth_weight[d];
        double wo = w_obs * og;
        wo = fmin(fmax(wo, -prm.clip_grad_scale), prm.clip_grad_scale);
        const double ws = w_smooth * sg;
        const double gt = wo + ws;
        s_grad[k] = gt;
        red7[0] += wo * wo; red7[1] += ws * ws; red7[2] += gt * gt;
        // check_joint_limit (optimizer.py:166-174, sic: scalar "any below" times elementwise "above")
        if (xc < a.rp.lower[d] - 5e-3) red7[5] = 1.0;
        if (xc > a.rp.upper[d] + 5e-3) red7[6] = 1.0;
        if (goal_set && i == n - 1) { const double dg = xc - s_end[d]; red7[4] += dg * dg; }
    }
    for (int k = tid; k < (n + 1) * ND; k += nthr) {   // smoothness loss rows 0..n (cost.py:443-445)
        const int r = k / ND, d = k - r * ND;
        double v;
        if (r == 0) v = (s_xi[d] - s_start[d]) * inv_dt;
        else if (r < n) v = (s_xi[k] - s_xi[k - ND]) * inv_dt;
        else v = goal_set ? 0.0 : (s_end[d] - s_xi[(n - 1) * ND + d]) * inv_dt;
        v *= prm.link_smooth_weight[d];
        red7[3] += v * v;
    }
    block_sum_n<7>(red7, s_red, bs_flip);
    const double norm_wo = sqrt(red7[0]), norm_ws = sqrt(red7[1]), norm_g = sqrt(red7[2]);
    const double smooth_sum = 0.5 * red7[3];
    const double goal_dist = goal_set ? sqrt(red7[4]) : 0.0;
    const bool violate = (red7[5] > 0.0) && (red7[6] > 0.0);
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
        OMGB_PROF(9);
        // ---- phase 6: covariant update -----------------------------------------------------------
        metric_apply(a, n, s_grad, s_u, s_viol);   // (s_viol .. s_viol + 2*n*9: scratch, see make_layout)
        for (int k = tid; k < n * ND; k += nthr) {   // + Trajectory.update (core.py:43-51)
            const int i = k / ND, d = k - i * ND;
            double v = s_xi[k];
            if (d < 7 || prm.consider_finger) {   // cfg.consider_finger: the finger DOFs move too (core.py:47-48)
                double up = -step_size * s_u[k];
                if (goal_set) {
                    double t1 = 0.0, t2 = 0.0;
                    for (int r = 0; r < c; ++r) {
                        const double m = __ldg(a.proj + (size_t)i * c + r);
                        t1 = fma(m, s_u[(n - c + r) * ND + d], t1);
                        t2 = fma(m, s_xi[(n - c + r) * ND + d] - s_goal[r * ND + d], t2);
                    }
                    up = up + step_size * t1 - t2;
                }
                v += up;
            }
            if (d >= 7) v = fmin(fmax(v, 0.0), 0.04);   // core.py:51
            s_viol[k] = v;   // staged: other threads still read s_xi rows n-c..n-1
        }
        __syncthreads();
        for (int k = tid; k < n * ND; k += nthr) s_xi[k] = s_viol[k];
        __syncthreads();
        OMGB_PROF(10);
        // ---- phase 7: smooth joint-limit projection (optimizer.py:148-164) ------------------------
        for (int round = 0; round <= prm.joint_limit_max_steps; ++round) {
            double t[1] = {0.0};
            double bm = -1.0;
            int bk = 0x7fffffff;
            for (int k = tid; k < n * ND; k += nthr) {
                const int d = k % ND;
                const double v = s_xi[k];
                double viol = 0.0;
                if (v < a.rp.lower[d]) viol = a.rp.lower[d] - v;
                else if (v > a.rp.upper[d]) viol = a.rp.upper[d] - v;
                s_viol[k] = viol;
                t[0] += viol * viol;
                const double m = fabs(viol);
                if (m > bm) { bm = m; bk = k; }   // first occurrence of the max in flat order (np.argmax)
            }
            block_sum_n<1>(t, s_red, bs_flip);
            const double vn = sqrt(t[0]);
            if (!(vn > 1e-2) || round == prm.joint_limit_max_steps) break;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double om = __shfl_xor_sync(0xffffffffu, bm, off);
                const int ok = __shfl_xor_sync(0xffffffffu, bk, off);
                if (om > bm || (om == bm && ok < bk)) { bm = om; bk = ok; }
            }
            double *s_am = s_red + L.red_max;   // per-warp maxima, then the block's
            if (lane == 0) { s_am[warp] = bm; s_hist[warp] = bk; }
            __syncthreads();
            if (tid == 0) {
                double fm = s_am[0];
                int fk = s_hist[0];
                for (int w = 1; w < nwarps; ++w)
                    if (s_am[w] > fm || (s_am[w] == fm && s_hist[w] < fk)) { fm = s_am[w]; fk = s_hist[w]; }
                s_am[nwarps] = fm;
                s_hist[260] = fk;
            }
            __syncthreads();
            const double vmax = s_am[nwarps];
            const int kmax = s_hist[260];
            metric_apply(a, n, s_viol, s_u, s_viol + n * ND);
            const double scale = vmax / (fabs(s_u[kmax]) + 1e-8);
            for (int k = tid; k < n * ND; k += nthr) s_xi[k] += scale * s_u[k];
            ++limit_rounds;
            __syncthreads();
        }
        // ---- phase 8: write back ---------------------------------------------------------------
        for (int k = tid; k < n * ND; k += nthr) g_xi[k] = s_xi[k];
    }
    OMGB_PROF(11);
    if (a.prof && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.prof[(size_t)b * 16 + 15] = (long long)gt;
    }
    if (tid == 0) {
        double *inf = s_red + L.red_max;   // staged in shared memory, written as one coalesced 128-byte row below
        inf[OMGB_INFO_OBS] = obs_sum;
        inf[OMGB_INFO_SMOOTH] = smooth_sum;
        inf[OMGB_INFO_COST] = w_obs * obs_sum + w_smooth * smooth_sum;
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
        inf[OMGB_INFO_RESERVED] = (double)n_act;   // link instances that survived the cull (diagnostic)
        if (a.done && a.stop_on_terminate && terminate && iteration > 0) a.done[b] = 1;
        if (a.cta_cost) a.cta_cost[b] = (int)min((long long)0x7fffffff, clock64() - t_begin);
    }
    __syncwarp();
    if (tid < OMGB_INFO_STRIDE) {
        const double v = s_red[L.red_max + tid];
        a.info[(size_t)b * OMGB_INFO_STRIDE + tid] = v;
        if (a.hist_info) a.hist_info[((size_t)iteration * a.batch + b) * OMGB_INFO_STRIDE + tid] = v;
    }
    if (a.hist_xi) {   // planner.py:622: history_trajectories.append(np.copy(traj.data))
        double *h = a.hist_xi + ((size_t)iteration * a.batch + b) * (size_t)(n * ND);
        for (int k = tid; k < n * ND; k += nthr) h[k] = s_xi[k];
    }
}

// A trajectory frozen by stop_on_terminate keeps its last state and info row in the history (the reference's loop has
// ended for it, planner.py:627-628; the host mirror truncates at the recorded iteration).
__device__ __forceinline__ void copy_history(const StepArgs &a, int b, int it) {
    if (it <= 0) return;
    const int n = a.prm.n_waypoints;
    if (a.hist_xi) {
        const double *src = a.hist_xi + ((size_t)(it - 1) * a.batch + b) * (size_t)(n * ND);
        double *dst = a.hist_xi + ((size_t)it * a.batch + b) * (size_t)(n * ND);
        for (int k = threadIdx.x; k < n * ND; k += blockDim.x) dst[k] = __ldcg(src + k);
    }
    if (a.hist_info && threadIdx.x < OMGB_INFO_STRIDE)
        a.hist_info[((size_t)it * a.batch + b) * OMGB_INFO_STRIDE + threadIdx.x] =
            __ldcg(a.hist_info + ((size_t)(it - 1) * a.batch + b) * OMGB_INFO_STRIDE + threadIdx.x);
}

// One launch = one iteration of every trajectory, CTA per trajectory (Optimizer.optimize granularity).
template <int LPI, int THREADS, int MINB, bool TOPK>
__global__ void __launch_bounds__(THREADS, MINB) chomp_step_kernel(const StepArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    if ((int)blockIdx.x >= a.batch) return;
    const int b = a.order ? a.order[blockIdx.x] : (int)blockIdx.x;
    if ((a.active && !a.active[b]) || (a.done && a.done[b])) {
        if (a.cta_cost && threadIdx.x == 0) a.cta_cost[b] = 0;
        if (a.done && a.done[b]) copy_history(a, b, a.iteration);
        return;
    }
    unsigned bulk_uses = 0u;
    chomp_iteration<LPI, TOPK>(a, smem, b, a.prm.obstacle_weight, a.prm.smoothness_weight, a.prm.step_size, a.iteration,
                               bulk_uses);
}

// Persistent plan kernel: ONE launch runs `iters` iterations of every trajectory (the fixed-goal inner loop of
// Planner.plan, omg/planner.py:612-627).  The grid is one CTA per resident slot; CTAs pull (iteration, trajectory)
// items from a global counter in iteration-major order.  Item (it, b) depends only on item (it - 1, b), which has a
// smaller id and therefore was claimed earlier by a CTA that is running: waiting on progress[b] cannot deadlock.
// Trajectories are independent, so there is no grid-wide barrier between iterations and the machine stays full
// until the last items of the whole plan (a per-iteration launch drains to a tail every iteration).
struct PlanArgs {
    const double *sched;     // [iters][3]: obstacle weight, smoothness weight, step size per iteration
    int *progress;           // [B]: iterations of trajectory b completed so far
    unsigned *counter;       // next item
    int iters;
};

template <int LPI, int THREADS, int MINB, bool TOPK>
__global__ void __launch_bounds__(THREADS, MINB) chomp_plan_kernel(const StepArgs a, const PlanArgs p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned s_item;
    const unsigned total = (unsigned)a.batch * (unsigned)p.iters;
    unsigned bulk_uses = 0u;   // bulk stagings this CTA has waited for (phase of the staging mbarrier)
    for (;;) {
        // every thread's generic-proxy accesses to the staging destinations, then the barrier, then the next item's
        // bulk copies (async proxy) into the same shared memory
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();   // (the previous item's shared memory is dead)
        if (threadIdx.x == 0) s_item = atomicAdd(p.counter, 1u);
        __syncthreads();
        const unsigned item = s_item;
        if (item >= total) break;
        const int it = (int)(item / (unsigned)a.batch), b = (int)(item % (unsigned)a.batch);
        if (threadIdx.x == 0) {
            volatile int *pr = p.progress + b;
            while (*pr < it) __nanosleep(200);
            __threadfence();   // acquire: the other CTA's xi / done writes are visible below
        }
        __syncthreads();
        const bool skip = a.done && __ldcg(a.done + b);   // frozen by stop_on_terminate
        if (!skip)
            chomp_iteration<LPI, TOPK>(a, smem, b, __ldg(p.sched + 3 * it), __ldg(p.sched + 3 * it + 1),
                                       __ldg(p.sched + 3 * it + 2), it, bulk_uses);
        else
            copy_history(a, b, it);
        __syncthreads();       // every thread's global stores of this item are issued ...
        if (threadIdx.x == 0) {
            __threadfence();   // ... and visible device-wide before the trajectory is released
            atomicExch(p.progress + b, it + 1);
        }
    }
}

}  // namespace omgb
