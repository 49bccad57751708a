// Goal scoring for the online goal-set learner, fused: the device half of Learner.cost_vector
// (omg/online_learner.py:104-150).  For every (trajectory b, goal g) one CTA
//   1. interpolates n' configurations on the joint-space line from the trajectory's current waypoint to the goal
//      (omg/util.py:261-290, mode "linear"),
//   2. runs the Panda forward kinematics of those n' + 1 configurations (robot_pykdl.py:148-215),
//   3. samples the obstacle potential of every body point (layers/sdf_matching_loss_kernel.cu:97-181, value only)
//      and weights it by the point's workspace speed (omg/cost.py:235-275, omg/config.py:162-187),
//   4. reduces over waypoints, links and body points to ONE number (online_learner.py:147-150).
// The reference materialises [G*n', 10, p] potentials AND gradients AND collision flags through three torch tensors
// and 4 kernel launches, then sums on the host side; here nothing but the [B, G] cost matrix leaves the SM.
#pragma once
#include "chomp_kernels.cuh"

namespace omgb {

struct GoalArgs {
    const ObjRec *objs;
    const float *grids;
    QuadDesc quad;
    const RobotConst *robot;
    const double *from;        // [B] rows of 9, row stride from_stride doubles (traj.data[start] of every trajectory)
    long long from_stride;
    const double *goals;       // [B,G,9], or [G,9] when goal_stride_b == 0
    long long goal_stride_b;
    float *costs;              // [B,G]
    DilDesc dil;
    RobotParams rp;
    int num_objects, num_goals, arc, finger_soft;
    float inv_dt;
    unsigned off_q, off_sc, off_frames, off_mask, off_part, off_act, off_objs, smem_total;
};

__host__ inline void goal_layout(GoalArgs &a) {
    const int cfgs = a.arc + 1;
    unsigned o = 0;
    a.off_q = o; o += sizeof(double) * cfgs * ND;
    o = align_up(o, 16);
    a.off_sc = o; o += sizeof(double2) * cfgs * 7;
    a.off_frames = o; o += sizeof(double) * cfgs * NL * 12;
    a.off_mask = o; o += sizeof(unsigned long long) * a.arc * NL;
    a.off_part = o; o += sizeof(double) * (a.arc * NL + 32);
    a.off_act = o; o += align_up(sizeof(unsigned short) * (a.arc * NL + 2) + 8, 8);
    o = align_up(o, 16);
    a.off_objs = o; o += sizeof(ObjRec) * a.num_objects;
    a.smem_total = align_up(o, 16);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) goal_cost_kernel(const GoalArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *s_q = reinterpret_cast<double *>(smem + a.off_q);
    double2 *s_sc = reinterpret_cast<double2 *>(smem + a.off_sc);
    double *s_frames = reinterpret_cast<double *>(smem + a.off_frames);
    unsigned long long *s_mask = reinterpret_cast<unsigned long long *>(smem + a.off_mask);
    double *s_part = reinterpret_cast<double *>(smem + a.off_part);
    int *s_count = reinterpret_cast<int *>(smem + a.off_act);                       // active link instances
    unsigned short *s_act = reinterpret_cast<unsigned short *>(smem + a.off_act + 8);
    ObjRec *s_objs = reinterpret_cast<ObjRec *>(smem + a.off_objs);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = THREADS / 32;
    const int g = blockIdx.x % a.num_goals, b = blockIdx.x / a.num_goals;
    const int arc = a.arc, cfgs = arc + 1, O = a.num_objects;
    const RobotConst *__restrict__ rc = a.robot;
    const int P = rc->p;

    // ---- stage the object records; interpolate (slot 0 = the trajectory's waypoint, slots 1..arc = the line) ----
    {
        const int words = (int)(sizeof(ObjRec) / 4) * O;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.objs);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_objs);
        for (int k = tid; k < words; k += THREADS) dst[k] = src[k];
        if (tid == 0) *s_count = 0;
    }
    {
        const double *qf = a.from + (size_t)b * a.from_stride;
        const double *qg = a.goals + (size_t)b * a.goal_stride_b + (size_t)g * ND;
        const double step = 1.0 / (double)(arc + 1);   // np.linspace(0, 1, arc + 2): i * step
        for (int k = tid; k < cfgs * ND; k += THREADS) {
            const int i = k / ND, d = k - i * ND;
            double v = qf[d];
            if (i > 0) {
                // scipy interp1d(kind="linear") on the two knots (0, from), (1, goal): w_hi * y_hi + w_lo * y_lo
                const double t = (double)i * step;
                v = __dadd_rn(__dmul_rn(t, qg[d]), __dmul_rn(__dsub_rn(1.0, t), v));
            }
            s_q[k] = v;
        }
    }
    __syncthreads();
    // ---- forward kinematics: sin/cos table, then 3 threads (one transform row each) per configuration ----
    for (int k = tid; k < cfgs * 7; k += THREADS) {
        const int cfg = k / 7, i = k - cfg * 7;
        double sn, cs;
        sincos(s_q[cfg * ND + i], &sn, &cs);
        s_sc[k] = make_double2(sn, cs);
    }
    __syncthreads();
    for (int k = tid; k < cfgs * 3; k += THREADS) {
        const int cfg = k / 3, r = k - cfg * 3;
        panda_fk_row(a.rp, s_q + cfg * ND, s_sc + cfg * 7, r, s_frames + (size_t)cfg * NL * 12);
    }
    __syncthreads();
    // ---- cull every (waypoint, link) bounding sphere against every object (same tests as the fused step) ----
    const int n_li = arc * NL;
    const bool use_dil = a.dil.enabled != 0;
    for (int base_li = 0; base_li < n_li; base_li += THREADS) {   // (uniform trip count: the ballot below needs whole warps)
        const int li = base_li + tid;
        unsigned long long m = 0ull;
        if (li < n_li) {
        const int j = li % NL;
        double cx, cy, cz;
        xform(s_frames + (size_t)(li + NL) * 12, (double)a.rp.sph[j][0], (double)a.rp.sph[j][1], (double)a.rp.sph[j][2],
              cx, cy, cz);
        const float fx = (float)cx, fy = (float)cy, fz = (float)cz, rad = a.rp.sph[j][3];
        for (int o = 0; o < O; ++o) {
            const ObjRec &ob = s_objs[o];
            if (ob.dis > 0.0f) continue;
            {   // first level: link sphere vs the object's world-frame bounding sphere (most pairs end here)
                const float dx = fx - ob.wsx, dy = fy - ob.wsy, dz = fz - ob.wsz, rr = rad + ob.wsr;
                if (ob.wsr >= 0.0f && dx * dx + dy * dy + dz * dz > rr * rr) continue;
            }
            const float qx = ob.r[0] * fx + ob.r[1] * fy + ob.r[2] * fz + ob.tx;
            const float qy = ob.r[3] * fx + ob.r[4] * fy + ob.r[5] * fz + ob.ty;
            const float qz = ob.r[6] * fx + ob.r[7] * fy + ob.r[8] * fz + ob.tz;
            const float s = rad + ob.cull_pad;
            const bool box = (qx > ob.lox - s) & (qx < ob.hix + s) & (qy > ob.loy - s) & (qy < ob.hiy + s) &
                             (qz > ob.loz - s) & (qz < ob.hiz + s);
            if (!box) continue;
            if (use_dil && ob.cull_pad < 1e29f) {
                const bool miss = (qx + s < ob.alox) | (qx - s > ob.ahix) | (qy + s < ob.aloy) | (qy - s > ob.ahiy) |
                                  (qz + s < ob.aloz) | (qz - s > ob.ahiz);
                if (miss) {
                    const bool interior =
                        ((qx - s - ob.minx) * ob.isx >= 1.5f) & ((qx + s - ob.minx) * ob.isx <= ob.fd0 - 1.5f) &
                        ((qy - s - ob.miny) * ob.isy >= 1.5f) & ((qy + s - ob.miny) * ob.isy <= ob.fd1 - 1.5f) &
                        ((qz - s - ob.minz) * ob.isz >= 1.5f) & ((qz + s - ob.minz) * ob.isz <= ob.fd2 - 1.5f);
                    if (interior) continue;
                }
            }
            m |= 1ull << o;
        }
        s_mask[li] = m;
        s_part[li] = 0.0;
        }
        // compaction of the instances that still have work (order is irrelevant: every instance owns its s_part slot)
        const unsigned bal = __ballot_sync(0xffffffffu, m != 0ull);
        int wbase = 0;
        if (lane == 0 && bal) wbase = atomicAdd(s_count, __popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (m != 0ull) s_act[wbase + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)li;
    }
    __syncthreads();
    const int n_act = *s_count;
    // ---- body points: half-warp per (waypoint, link), lane per body point (p <= 16) or warp per instance ----
    const int LPI = P <= 16 ? 16 : 32;
    const int gpw = 32 / LPI;
    const int sub = lane / LPI, pl = lane % LPI;
    const unsigned gmask = (LPI == 32) ? 0xffffffffu : (0xffffu << (sub * 16));
    for (int base = warp * gpw; base < n_act; base += NWARPS * gpw) {
        const int idx = base + sub;
        const bool have = idx < n_act;
        const int li = have ? (int)s_act[idx] : 0;
        unsigned long long m = have ? s_mask[li] : 0ull;
        const bool live = have && pl < P;
        if (!live) m = 0ull;
        const int i = have ? li / NL : 0, j = have ? li - i * NL : 0;
        const double *F = s_frames + (size_t)(li + NL) * 12;
        const double *bp = rc->pts[j][pl < P ? pl : 0];
        double X, Y, Z;
        xform(F, bp[0], bp[1], bp[2], X, Y, Z);
        const float x = (float)X, y = (float)Y, z = (float)Z;   // omg/cost.py:218 .float()
        float pot = 0.0f;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            unsigned mm = half ? (unsigned)(m >> 32) : (unsigned)m;
            while (mm) {
                const int o = __ffs(mm) - 1 + 32 * half;
                mm &= mm - 1;
                if (use_dil && classify_pair(s_objs[o], a.dil, o, x, y, z) != PAIR_EXACT) continue;
                float po, co;
                pair_potential(s_objs[o], a.grids, a.quad, x, y, z, po, co);
                pot = __fadd_rn(pot, po);
            }
        }
        if (a.finger_soft && j >= 8) pot = __fmul_rn(pot, 0.1f);   // omg/cost.py:350-353
        float val = 0.0f;
        if (pot != 0.0f) {
            // workspace speed against the previous configuration of the line (slot i; slot 0 = traj.data[start])
            const double *Fp = F - NL * 12;
            double Xp, Yp, Zp;
            xform(Fp, bp[0], bp[1], bp[2], Xp, Yp, Zp);
            const float vx = (x - (float)Xp) * a.inv_dt, vy = (y - (float)Yp) * a.inv_dt, vz = (z - (float)Zp) * a.inv_dt;
            val = pot * sqrtf(vx * vx + vy * vy + vz * vz);
        }
        double acc = (double)val;
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) acc += __shfl_xor_sync(gmask, acc, off, 16);
        if (LPI == 32) acc += __shfl_xor_sync(gmask, acc, 16, 32);
        if (have && pl == 0) s_part[li] = acc;
    }
    __syncthreads();
    // ---- deterministic reduction over the link instances ----
    if (warp == 0) {
        double t = 0.0;
        for (int k = lane; k < n_li; k += 32) t += s_part[k];
        t = warp_sum(t);
        if (lane == 0) a.costs[(size_t)b * a.num_goals + g] = (float)t;
    }
}

}  // namespace omgb
