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
//
// Step 3 runs in two stages per warp.  Stage A classifies every surviving (body point, object) pair against the
// lower-bound grid (one load proves "farther than eps"); the few pairs that need the operator -- ~7 % of the in-bounds
// pairs, scattered over the lanes -- are pushed into a warp-private queue in shared memory (point, speed, object).
// Stage B evaluates queued pairs 32 at a time, every lane busy: run in place, the operator's ~120 instructions were
// issued for warps with one or two live lanes and made up 28 % of the kernel's instructions
// (profiles/r02p_ncu_goal_cost_lines.txt).  The sum is order-free: every pair's potential x speed is an exact fp64
// product of two floats, accumulated in 2^-40 fixed point (int64), so the result does not depend on queue order,
// block shape or SDF layout.  (The reference's own sum over objects is an atomicAdd in arbitrary order followed by a
// torch fp32 tree reduction: its value is defined to ~1e-6 relative; tests/test_gpu_goal_scoring.py holds 2e-5.)
//
// The lines get shorter as the plan proceeds (n' = timesteps - start, online_learner.py:109-114): a CTA takes as many
// goals of its trajectory as fit ~30 configurations (1 goal at n' = 30, 6 at n' = 5), sharing the trajectory's
// waypoint (slot 0) and its forward kinematics, so that late iterations do not pay a CTA's fixed cost per 5-waypoint
// line.
#pragma once
#include "chomp_kernels.cuh"

namespace omgb {

constexpr int GOAL_QCAP = 64;                       // queue entries per warp (a push adds <= 32 to <= 31 pending)
constexpr double GOAL_FIX = 1099511627776.0;        // 2^40
constexpr int GOAL_MAX_GPC = 32;                    // goals per CTA

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
    int gpc, ctas_per_traj;    // goals per CTA; CTAs per trajectory = ceil(num_goals / gpc)
    float inv_dt;
    unsigned off_q, off_sc, off_queue, off_qobj, off_frames, off_mask, off_mask_hi, off_act, off_red, off_objs,
        off_sph, smem_total;
};

__host__ inline void goal_layout(GoalArgs &a, int warps) {
    const int cfgs = a.gpc * a.arc + 1, n_li = a.gpc * a.arc * NL;
    unsigned o = 0;
    // region A: joint values + sin/cos table (dead once the frames exist), reused by the per-warp queues
    a.off_q = 0;
    a.off_sc = align_up((unsigned)(sizeof(double) * cfgs * ND), 16);
    const unsigned fk_bytes = a.off_sc + (unsigned)(sizeof(double2) * cfgs * 7);
    a.off_queue = 0;
    a.off_qobj = (unsigned)(sizeof(float4) * GOAL_QCAP * warps);
    const unsigned q_bytes = a.off_qobj + (unsigned)(sizeof(unsigned short) * GOAL_QCAP * warps);
    o = align_up(fk_bytes > q_bytes ? fk_bytes : q_bytes, 16);
    a.off_frames = o; o += sizeof(double) * cfgs * NL * 12;
    a.off_mask = o; o += sizeof(unsigned) * n_li;
    a.off_mask_hi = o; if (a.num_objects > 32) o += sizeof(unsigned) * n_li;
    a.off_act = o; o += align_up((unsigned)(sizeof(unsigned short) * (n_li + 2)) + 8, 8);
    a.off_red = o; o += sizeof(unsigned long long) * GOAL_MAX_GPC;
    o = align_up(o, 16);
    a.off_objs = o; o += sizeof(ObjRec) * a.num_objects;
    a.off_sph = o; o += sizeof(float4) * a.num_objects;
    a.smem_total = align_up(o, 16);
}

// LPI: lanes per link instance (16 when the robot has <= 16 body points per link, else 32); HI: more than 32 objects.
template <int THREADS, int LPI, bool HI>
__global__ void __launch_bounds__(THREADS, (THREADS <= 128 ? 8 : THREADS <= 192 ? 6 : THREADS <= 256 ? 5 : 4))
goal_cost_kernel(const GoalArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *s_q = reinterpret_cast<double *>(smem + a.off_q);
    double2 *s_sc = reinterpret_cast<double2 *>(smem + a.off_sc);
    double *s_frames = reinterpret_cast<double *>(smem + a.off_frames);
    unsigned *s_mask = reinterpret_cast<unsigned *>(smem + a.off_mask);
    unsigned *s_mask_hi = reinterpret_cast<unsigned *>(smem + a.off_mask_hi);
    int *s_count = reinterpret_cast<int *>(smem + a.off_act);                       // active link instances
    unsigned short *s_act = reinterpret_cast<unsigned short *>(smem + a.off_act + 8);   // (config << 4) | link
    unsigned long long *s_cost = reinterpret_cast<unsigned long long *>(smem + a.off_red);   // per goal of this CTA
    ObjRec *s_objs = reinterpret_cast<ObjRec *>(smem + a.off_objs);
    float4 *s_sph = reinterpret_cast<float4 *>(smem + a.off_sph);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = THREADS / 32;
    const int b = blockIdx.x / a.ctas_per_traj, g0 = (blockIdx.x % a.ctas_per_traj) * a.gpc;
    const int ng = min(a.gpc, a.num_goals - g0);          // goals of this CTA: g0 .. g0 + ng - 1
    // configuration slots: 0 = the trajectory's waypoint, then `arc` per goal line
    const int arc = a.arc, cfgs = ng * arc + 1, O = a.num_objects;
    const RobotConst *__restrict__ rc = a.robot;
    const int P = rc->p;

    // ---- stage the object records; interpolate (slot 0 = the trajectory's waypoint, slots 1..arc = the line) ----
    {
        const int quads = (int)(sizeof(ObjRec) / 16) * O;
        const uint4 *src = reinterpret_cast<const uint4 *>(a.objs);
        uint4 *dst = reinterpret_cast<uint4 *>(s_objs);
        for (int k = tid; k < quads; k += THREADS) dst[k] = __ldg(src + k);
        if (tid == 0) *s_count = 0;
        if (tid < GOAL_MAX_GPC) s_cost[tid] = 0ull;
    }
    {
        const double *qf = a.from + (size_t)b * a.from_stride;
        const double *qg0 = a.goals + (size_t)b * a.goal_stride_b + (size_t)g0 * ND;
        const double step = 1.0 / (double)(arc + 1);   // np.linspace(0, 1, arc + 2): i * step
        for (int k = tid; k < cfgs * ND; k += THREADS) {
            const int sl = k / ND, d = k - sl * ND;
            double v = qf[d];
            if (sl > 0) {
                // scipy interp1d(kind="linear") on the two knots (0, from), (1, goal): w_hi * y_hi + w_lo * y_lo
                const int l = (sl - 1) / arc, i = sl - l * arc;
                const double t = (double)i * step;
                v = __dadd_rn(__dmul_rn(t, qg0[(size_t)l * ND + d]), __dmul_rn(__dsub_rn(1.0, t), v));
            }
            s_q[k] = v;
        }
    }
    __syncthreads();
    sph_stage(s_objs, s_sph, O);   // (read after the barriers of the FK phase)
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
    const int n_li = ng * arc * NL;
    const bool use_dil = a.dil.enabled != 0;
    for (int base_li = 0; base_li < n_li; base_li += THREADS) {   // (uniform trip count: the ballot below needs whole warps)
        const int li = base_li + tid;
        unsigned mlo = 0u, mhi = 0u;
        int ci = 0, cj = 0;
        if (li < n_li) {
            ci = li / NL; cj = li - ci * NL;
            double cx, cy, cz;
            xform(s_frames + (size_t)(li + NL) * 12, (double)a.rp.sph[cj][0], (double)a.rp.sph[cj][1],
                  (double)a.rp.sph[cj][2], cx, cy, cz);
            const float fx = (float)cx, fy = (float)cy, fz = (float)cz, rad = a.rp.sph[cj][3];
            // first level: link sphere vs the objects' world-frame bounding spheres (most pairs end here)
            unsigned long long near = sph_near(s_sph, 0, O < 32 ? O : 32, fx, fy, fz, rad);
            if (HI) near |= (unsigned long long)sph_near(s_sph, 32, O, fx, fy, fz, rad) << 32;
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
                if (!HI || o < 32) mlo |= 1u << o;
                else mhi |= 1u << (o - 32);
            }
            s_mask[li] = mlo;
            if (HI) s_mask_hi[li] = mhi;
        }
        // compaction of the instances that still have work (order is irrelevant: the sum is order-free)
        const bool any = (mlo | mhi) != 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, any);
        int wbase = 0;
        if (lane == 0 && bal) wbase = atomicAdd(s_count, __popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (any) s_act[wbase + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)((ci << 4) | cj);
    }
    __syncthreads();   // (the sin/cos table and the joint values are dead: region A now holds the queues)
    const int n_act = *s_count;
    float4 *q_pt = reinterpret_cast<float4 *>(smem + a.off_queue) + warp * GOAL_QCAP;   // x, y, z, speed
    unsigned short *q_ob = reinterpret_cast<unsigned short *>(smem + a.off_qobj) + warp * GOAL_QCAP;   // object | finger flag | goal << 8
    int q_cnt = 0;
    // stage B for queue slot k (k < 0: this lane has none): the operator's value-only evaluation, potential x speed in
    // exact fp64, fixed point; then one integer add per goal present in the batch (redux over 21-bit limbs)
    auto evaluate = [&](int k) {
        unsigned long long f = 0ull;
        unsigned l = 0xffffffffu;
        if (k >= 0) {
            const float4 it = q_pt[k];
            const unsigned oo = q_ob[k];
            float po, co;
            pair_potential(s_objs[oo & 63u], a.grids, a.quad, it.x, it.y, it.z, po, co);
            double v = (double)po * (double)it.w;
            if (oo & 0x80u) v *= (double)0.1f;   // omg/cost.py:350-353 (soft finger links)
            f = (unsigned long long)__double2ll_rn(v * GOAL_FIX);   // (v >= 0)
            if (f != 0ull) l = oo >> 8;
        }
        unsigned pending = __ballot_sync(0xffffffffu, l != 0xffffffffu);
        while (pending) {
            const unsigned l0 = __shfl_sync(0xffffffffu, l, __ffs(pending) - 1);
            const bool mine = l == l0;
            const unsigned long long w = mine ? f : 0ull;
            const unsigned s0 = __reduce_add_sync(0xffffffffu, (unsigned)(w & 0x1fffffull));
            const unsigned s1 = __reduce_add_sync(0xffffffffu, (unsigned)((w >> 21) & 0x1fffffull));
            const unsigned s2 = __reduce_add_sync(0xffffffffu, (unsigned)(w >> 42));
            if (lane == 0)
                atomicAdd(s_cost + l0, (unsigned long long)s0 + ((unsigned long long)s1 << 21) + ((unsigned long long)s2 << 42));
            pending &= ~__ballot_sync(0xffffffffu, mine);
        }
    };
    // ---- body points: LPI lanes per (waypoint, link), lane per body point ----
    constexpr int GPW = 32 / LPI;
    const int sub = lane / LPI, pl = lane % LPI;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int base = warp * GPW; base < n_act; base += NWARPS * GPW) {
        const int idx = base + sub;
        unsigned nlo = 0u, nhi = 0u;   // objects whose operator value this point needs
        float x = 0.0f, y = 0.0f, z = 0.0f, sp = 0.0f;
        unsigned tag = 0u;
        if (idx < n_act && pl < P) {
            const unsigned v = s_act[idx];
            const int i = (int)(v >> 4), j = (int)(v & 15u), li = i * NL + j;   // i: configuration index over the CTA's lines
            const double *F = s_frames + (size_t)(li + NL) * 12;
            const double *bp = rc->pts[j][pl];
            const double b0 = bp[0], b1 = bp[1], b2 = bp[2];
            double X, Y, Z;
            xform(F, b0, b1, b2, X, Y, Z);
            x = (float)X; y = (float)Y; z = (float)Z;   // omg/cost.py:218 .float()
            unsigned mm = s_mask[li];
            while (mm) {
                const int o = __ffs(mm) - 1;
                mm &= mm - 1;
                if (!use_dil || classify_pair(s_objs[o], a.dil, o, x, y, z) == PAIR_EXACT) nlo |= 1u << o;
            }
            if (HI) {
                mm = s_mask_hi[li];
                while (mm) {
                    const int o = __ffs(mm) - 1;
                    mm &= mm - 1;
                    if (!use_dil || classify_pair(s_objs[o + 32], a.dil, o + 32, x, y, z) == PAIR_EXACT) nhi |= 1u << o;
                }
            }
            if (nlo | nhi) {
                // workspace speed against the previous configuration of the line (a line's first one: slot 0 =
                // traj.data[start])
                const int l = i / arc;
                const double *Fp = (i - l * arc == 0) ? (s_frames + (size_t)j * 12) : (F - NL * 12);
                double Xp, Yp, Zp;
                xform(Fp, b0, b1, b2, Xp, Yp, Zp);
                const float vx = (x - (float)Xp) * a.inv_dt, vy = (y - (float)Yp) * a.inv_dt,
                            vz = (z - (float)Zp) * a.inv_dt;
                sp = sqrtf(vx * vx + vy * vy + vz * vz);
                tag = ((a.finger_soft && j >= 8) ? 0x80u : 0u) | ((unsigned)l << 8);
            }
        }
        // push this round's pairs, one object per lane and pass; evaluate whenever 32 are pending
        for (;;) {
            const bool has = (nlo | nhi) != 0u;
            const unsigned bal = __ballot_sync(0xffffffffu, has);
            if (!bal) break;
            if (has) {
                unsigned o;
                if (nlo) { o = (unsigned)(__ffs(nlo) - 1); nlo &= nlo - 1; }
                else { o = 32u + (unsigned)(__ffs(nhi) - 1); nhi &= nhi - 1; }
                const int pos = q_cnt + __popc(bal & lt_mask);
                q_pt[pos] = make_float4(x, y, z, sp);
                q_ob[pos] = (unsigned short)(o | tag);
            }
            q_cnt += __popc(bal);
            __syncwarp();
            if (q_cnt >= 32) {
                q_cnt -= 32;
                evaluate(q_cnt + lane);
                __syncwarp();
            }
        }
    }
    if (q_cnt > 0) evaluate(lane < q_cnt ? lane : -1);   // (warp-uniform)
    __syncthreads();
    if (tid < ng) a.costs[(size_t)b * a.num_goals + g0 + tid] = (float)((double)s_cost[tid] * (1.0 / GOAL_FIX));
}

}  // namespace omgb
