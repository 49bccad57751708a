// The online learner's goal re-weighting on the device (second half): what Learner.update_goal does on
// the host after the collision costs are known (omg/online_learner.py:151-160, 162-249), for every trajectory of a
// batch in one launch, so that a goal-set plan with goal switching never leaves the GPU between iterations:
//     omgb_goal_costs  ->  omgb_learner_update  ->  omgb_chomp_plan_step
//
//   cost vector   potentials = base_obstacle_weight * collision (fp32) + smoothness_base_weight * dist_eps * smooth,
//                 smooth = ||diff over the JOINT axis of (xi[first] - goal)||^2 (sic, :151-153), optionally divided by
//                 its L2 norm over the goals (:159-160)
//   update rule   FTL / FTC / Exp / MD / Proj (:177-235).  MD: five experts, each the Bregman projection onto the
//                 simplex with the shifted entropy (:32-58: <= 100 fixed-point rounds, each a <= 100-step bisection),
//                 then the expert mixture re-weighted INSIDE the expert loop like the reference (:230-235)
//   selection     goal_idx = argmax p (first maximum), end = goal_set[goal_idx], projection rows = reach_grasps[goal_idx]
//
// One CTA per trajectory, one warp per expert, lanes over goals (<= 256 goals).  fp64; sums are warp-shuffle trees, so
// results agree with numpy's pairwise sums to rounding (the bisection itself stops at 1e-6).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/omgb200.h"
#include "learner_bisect.h"

namespace omgb {

constexpr int LRN_EXPERTS = 5;
constexpr int LRN_MAX_GOALS = 256;
constexpr int LRN_THREADS = 32 * LRN_EXPERTS;

struct LearnerArgs {
    omgb_learner_params_t prm;
    const double *xi;            // [B,n,9]
    const float *collision;      // [B,G] (omgb_goal_costs); unused by Proj
    const double *goal_set;      // [B,G,9] or [G,9]
    long long goal_stride_b;
    const double *reach;         // [B,G,c,9] or [G,c,9]; null: rows = goal_set[idx]
    long long reach_stride_b;
    double *p, *sum_costs, *experts_p, *experts_costs, *q;   // state [B,G], [B,G], [B,5,G], [B,5], [B,5]
    const uint8_t *done;         // [B] or null
    int *goal_idx;               // [B] in/out
    double *end;                 // [B,9] out
    double *goal_rows;           // [B,c,9] out (c >= 1)
    double *cost_vector;         // [B,G] out or null
    int *selected_hist;          // [B] out or null (Planner.selected_goals slot of this iteration)
    int batch;
};

__device__ __forceinline__ double lrn_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double lrn_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// omg/online_learner.py:32-58 for one expert: x (previous distribution), v = eta * cv, both in shared memory [G];
// result written to out [G] (shared).  All 32 lanes of the calling warp participate; LRN_PER_LANE = goals per lane
// (a template parameter: 20 goals need one register set per lane, not eight -- occupancy is what hides the latency
// of the bisection's dependent exp / reduce chain).
// TWO_STEP: the bisection takes two steps per round (learner_bisect.h) -- same points, same comparisons, same result;
// chosen by the launcher for small batches, where the kernel's time is this warp's dependent chain.
template <int LRN_PER_LANE, bool TWO_STEP>
__device__ void lrn_bregman_warp(const double *x, const double *cv, double eta, int G, double delta, double *out) {
    const int lane = threadIdx.x & 31;
    double sh[LRN_PER_LANE], v[LRN_PER_LANE], alpha[LRN_PER_LANE], y[LRN_PER_LANE];
    double vmax = -1e300;
#pragma unroll
    for (int k = 0; k < LRN_PER_LANE; ++k) {
        const int g = lane + 32 * k;
        const bool in = g < G;
        sh[k] = in ? x[g] + delta : 0.0;
        v[k] = in ? eta * cv[g] : 0.0;
        alpha[k] = 0.0;
        y[k] = 0.0;
        if (in) vmax = fmax(vmax, 1.0 + v[k]);
    }
    const double x1 = lrn_warp_max(vmax);
    const double target = 1.0 + (double)G * delta;
    const double err = 1e-6;
    for (int it = 0; it < 100; ++it) {
        // find_zero (:18-30)
        double xx;
        if (TWO_STEP) {
            xx = lrn_find_zero_two(x1, err, [&](double x0, double st, double &f0, double &fm, double &fp) {
                const double xm = x0 - st, xp = x0 + st;
                double p0 = 0.0, pm = 0.0, pp = 0.0;
#pragma unroll
                for (int k = 0; k < LRN_PER_LANE; ++k)
                    if (lane + 32 * k < G) {
                        p0 += sh[k] * exp(x0 + (alpha[k] - v[k]));
                        pm += sh[k] * exp(xm + (alpha[k] - v[k]));
                        pp += sh[k] * exp(xp + (alpha[k] - v[k]));
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {   // three interleaved lrn_warp_sum trees
                    const double t0 = __shfl_xor_sync(0xffffffffu, p0, o), tm = __shfl_xor_sync(0xffffffffu, pm, o),
                                 tp = __shfl_xor_sync(0xffffffffu, pp, o);
                    p0 += t0; pm += tm; pp += tp;
                }
                f0 = p0 - target; fm = pm - target; fp = pp - target;
            });
        } else {
            xx = (0.0 + x1) / 2;
            double step = (x1 - 0.0) / 4;
            for (int k2 = 0; k2 < 100; ++k2) {
                double part = 0.0;
#pragma unroll
                for (int k = 0; k < LRN_PER_LANE; ++k)
                    if (lane + 32 * k < G) part += sh[k] * exp(xx + (alpha[k] - v[k]));
                const double f = lrn_warp_sum(part) - target;
                if (fabs(f) < err) break;
                const double sg = (f > 0.0) ? 1.0 : ((f < 0.0) ? -1.0 : 0.0);
                xx -= step * sg;
                step /= 2;
            }
        }
        const double lam = xx;
        double dn = 0.0, nxt[LRN_PER_LANE];
#pragma unroll
        for (int k = 0; k < LRN_PER_LANE; ++k) {
            nxt[k] = 0.0;
            if (lane + 32 * k < G) {
                y[k] = sh[k] * exp(lam + alpha[k] - v[k]) - delta;
                nxt[k] = fmax(0.0, v[k] - lam + log(delta / sh[k]));
                const double d = alpha[k] - nxt[k];
                dn += d * d;
            }
        }
        if (sqrt(lrn_warp_sum(dn)) < err) break;
#pragma unroll
        for (int k = 0; k < LRN_PER_LANE; ++k) alpha[k] = nxt[k];
    }
    double ys = 0.0;
#pragma unroll
    for (int k = 0; k < LRN_PER_LANE; ++k) {
        y[k] = fmax(y[k], 0.0);
        if (lane + 32 * k < G) ys += y[k];
    }
    ys = lrn_warp_sum(ys);
#pragma unroll
    for (int k = 0; k < LRN_PER_LANE; ++k)
        if (lane + 32 * k < G) out[lane + 32 * k] = y[k] / ys;
}

template <int PER_LANE, bool TWO_STEP = false>
__global__ void __launch_bounds__(LRN_THREADS, PER_LANE <= 2 ? 4 : 2) learner_update_kernel(const LearnerArgs a) {
    __shared__ double s_cv[LRN_MAX_GOALS];
    __shared__ double s_p[LRN_MAX_GOALS];
    __shared__ double s_old[LRN_EXPERTS][LRN_MAX_GOALS];
    __shared__ double s_new[LRN_EXPERTS][LRN_MAX_GOALS];
    __shared__ double s_red[LRN_EXPERTS + 3];
    __shared__ int s_idx;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const omgb_learner_params_t &P = a.prm;
    const int G = P.num_goals, n = P.n_waypoints, c = P.constraint_rows;
    if (a.done && a.done[b]) {   // the plan has ended for this trajectory: keep its goal
        if (a.selected_hist && tid == 0) a.selected_hist[b] = a.goal_idx[b];
        return;
    }
    const double *goals = a.goal_set + (size_t)b * a.goal_stride_b;
    const int alg = P.alg;

    // ---- cost vector (omg/online_learner.py:151-160) ----------------------------------------------------------
    if (alg != OMGB_LEARNER_PROJ) {
        const double *ts = a.xi + ((size_t)b * n + P.first_waypoint) * 9;
        for (int g = tid; g < G; g += LRN_THREADS) {
            double d[9], e[8];
#pragma unroll
            for (int k = 0; k < 9; ++k) d[k] = ts[k] - goals[(size_t)g * 9 + k];
#pragma unroll
            for (int k = 0; k < 8; ++k) { const double t = d[k + 1] - d[k]; e[k] = t * t; }
            const double ss = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));   // numpy's 8-wide sum
            const double nr = sqrt(ss);
            const double smooth = nr * nr;
            const float wc = (float)P.base_obstacle_weight * a.collision[(size_t)b * G + g];
            s_cv[g] = (double)wc + (P.smoothness_base_weight * P.dist_eps) * smooth;
        }
        __syncthreads();
        if (P.normalize_cost) {
            if (warp == 0) {
                double part = 0.0;
                for (int g = lane; g < G; g += 32) part += s_cv[g] * s_cv[g];
                part = lrn_warp_sum(part);
                if (lane == 0) s_red[0] = sqrt(part);
            }
            __syncthreads();
            const double nrm = s_red[0];
            for (int g = tid; g < G; g += LRN_THREADS) s_cv[g] = s_cv[g] / nrm;
            __syncthreads();
        }
        if (a.cost_vector)
            for (int g = tid; g < G; g += LRN_THREADS) a.cost_vector[(size_t)b * G + g] = s_cv[g];
    }

    // ---- update rule -> s_p ---------------------------------------------------------------------------------------
    double *p_g = a.p + (size_t)b * G;
    if (alg == OMGB_LEARNER_MD) {
        double *ep = a.experts_p + (size_t)b * LRN_EXPERTS * G;
        for (int k = tid; k < LRN_EXPERTS * G; k += LRN_THREADS) s_old[k / G][k % G] = ep[k];
        __syncthreads();
        const double delta = 1.0 / (double)(4 * G + 1);
        lrn_bregman_warp<PER_LANE, TWO_STEP>(s_old[warp], s_cv, P.etas[warp], G, delta, s_new[warp]);
        __syncwarp();
        {   // experts_costs[i] = cv . p + weights . |p - p_old| (:222-224), weights = 1
            double c1 = 0.0, c2 = 0.0;
            for (int g = lane; g < G; g += 32) {
                c1 += s_cv[g] * s_new[warp][g];
                c2 += fabs(s_new[warp][g] - s_old[warp][g]);
            }
            c1 = lrn_warp_sum(c1); c2 = lrn_warp_sum(c2);
            if (lane == 0) s_red[warp] = c1 + c2;
        }
        __syncthreads();
        if (tid == 0) {
            // the reference re-weights q after EVERY expert with the cost vector as it is at that moment: entries of
            // the experts not yet processed are still last iteration's (:228-229)
            double ec[LRN_EXPERTS], q[LRN_EXPERTS];
            for (int k = 0; k < LRN_EXPERTS; ++k) { ec[k] = a.experts_costs[(size_t)b * LRN_EXPERTS + k]; q[k] = a.q[(size_t)b * LRN_EXPERTS + k]; }
            for (int i = 0; i < LRN_EXPERTS; ++i) {
                ec[i] = s_red[i];
                double qs = 0.0;
                for (int k = 0; k < LRN_EXPERTS; ++k) { q[k] = q[k] * exp(-1.0 * ec[k]); qs += q[k]; }
                for (int k = 0; k < LRN_EXPERTS; ++k) q[k] = q[k] / qs;
            }
            for (int k = 0; k < LRN_EXPERTS; ++k) {
                a.experts_costs[(size_t)b * LRN_EXPERTS + k] = ec[k];
                a.q[(size_t)b * LRN_EXPERTS + k] = q[k];
                s_red[k] = q[k];
            }
        }
        __syncthreads();
        for (int k = tid; k < LRN_EXPERTS * G; k += LRN_THREADS) ep[k] = s_new[k / G][k % G];
        // p = sum_k experts_p[k] q[k], normalised (:232-235; only the last pass of the loop survives)
        for (int g = tid; g < G; g += LRN_THREADS) {
            double acc = 0.0;
            for (int k = 0; k < LRN_EXPERTS; ++k) acc += s_new[k][g] * s_red[k];
            s_p[g] = acc;
        }
        __syncthreads();
        if (warp == 0) {
            double part = 0.0;
            for (int g = lane; g < G; g += 32) part += s_p[g];
            part = lrn_warp_sum(part);
            if (lane == 0) s_red[LRN_EXPERTS] = part;
        }
        __syncthreads();
        for (int g = tid; g < G; g += LRN_THREADS) s_p[g] = s_p[g] / s_red[LRN_EXPERTS];
        __syncthreads();
    } else if (alg == OMGB_LEARNER_EXP) {
        double *sc = a.sum_costs + (size_t)b * G;
        for (int g = tid; g < G; g += LRN_THREADS) { sc[g] = sc[g] + s_cv[g]; s_old[0][g] = sc[g]; }
        __syncthreads();
        if (warp == 0) {
            double part = 0.0;
            for (int g = lane; g < G; g += 32) part += s_old[0][g];
            part = lrn_warp_sum(part);
            if (lane == 0) s_red[0] = part;
        }
        __syncthreads();
        for (int g = tid; g < G; g += LRN_THREADS) {
            const double norm_sum = s_old[0][g] / (s_red[0] + 1e-8);
            const double p_new = exp(-P.eta * s_cv[g]) * p_g[g];
            s_p[g] = p_new * 0.999 + norm_sum * 0.001;
        }
        __syncthreads();
        if (warp == 0) {
            double part = 0.0;
            for (int g = lane; g < G; g += 32) part += s_p[g];
            part = lrn_warp_sum(part);
            if (lane == 0) s_red[1] = part;
        }
        __syncthreads();
        for (int g = tid; g < G; g += LRN_THREADS) s_p[g] = s_p[g] / (s_red[1] + 1e-8);
        __syncthreads();
    } else {
        // FTL / FTC / Proj / INIT: a one-hot distribution at the argmin of some score (first minimum, like np.argmin)
        double *sc = a.sum_costs + (size_t)b * G;
        for (int g = tid; g < G; g += LRN_THREADS) {
            double score;
            if (alg == OMGB_LEARNER_FTL) { sc[g] = sc[g] + s_cv[g]; score = sc[g]; }
            else if (alg == OMGB_LEARNER_PROJ) {
                const double *last = a.xi + ((size_t)b * n + (n - 1)) * 9;
                double ss = 0.0;
                for (int k = 0; k < 9; ++k) { const double t = last[k] - goals[(size_t)g * 9 + k]; ss += t * t; }
                score = sqrt(ss);
            } else score = s_cv[g];
            s_old[0][g] = score;
        }
        __syncthreads();
        if (tid == 0) {
            int best = 0;
            for (int g = 1; g < G; ++g)
                if (s_old[0][g] < s_old[0][best]) best = g;
            s_idx = best;
        }
        __syncthreads();
        for (int g = tid; g < G; g += LRN_THREADS) s_p[g] = (g == s_idx) ? 1.0 : 0.0;
        __syncthreads();
    }

    // ---- selection (:237-249) ---------------------------------------------------------------------------------------
    if (alg != OMGB_LEARNER_INIT)
        for (int g = tid; g < G; g += LRN_THREADS) p_g[g] = s_p[g];
    if (tid == 0) {
        int best = 0;
        for (int g = 1; g < G; ++g)
            if (s_p[g] > s_p[best]) best = g;
        s_idx = best;
        a.goal_idx[b] = best;
        if (a.selected_hist) a.selected_hist[b] = best;
    }
    __syncthreads();
    const int idx = s_idx;
    if (tid < 9) a.end[(size_t)b * 9 + tid] = goals[(size_t)idx * 9 + tid];
    if (a.goal_rows) {
        if (a.reach) {
            const double *src = a.reach + (size_t)b * a.reach_stride_b + (size_t)idx * c * 9;
            for (int k = tid; k < c * 9; k += LRN_THREADS) a.goal_rows[(size_t)b * c * 9 + k] = src[k];
        } else {
            for (int k = tid; k < c * 9; k += LRN_THREADS) a.goal_rows[(size_t)b * c * 9 + k] = goals[(size_t)idx * 9 + (k % 9)];
        }
    }
}

}  // namespace omgb
