// libomgb200.so -- C ABI (include/omgb200.h) over the sm_100a kernels.  No torch, no Eigen, no Sophus.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/omgb200.h"
#include "host_common.h"
#include "chomp_kernels.cuh"
#include "goal_kernels.cuh"
#include "sdf_device.cuh"
#include "learner_kernels.cuh"
#include "sdf_asset_kernels.cuh"
#include "traj_kernels.cuh"

namespace omgb {

static thread_local std::string g_err;
static std::atomic<unsigned long long> g_launches{0};   // kernels launched by this library (bench.py's gpu_launches)
static std::mutex g_attr_mutex;                          // guards the per-instantiation function-attribute caches

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int host_fail(int code, const std::string &msg) { return fail(code, msg); }
void host_count_launch() { ++g_launches; }

#define OMGB_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(OMGB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// ----------------------------------------------------------------------------------------------------
// per-object preparation: quaternion of the pose (what Sophus::SE3<float>(Matrix4) holds,
// se3.hpp:387-389 / so3.hpp:392), its rotation matrix (so3().matrix(), kernel.cu:126), grid constants.
// All ops separately rounded (no contraction) so the record is bit-identical to oracle/sdf_loss_ref.c.
// ----------------------------------------------------------------------------------------------------
__global__ void prep_objects_kernel(const float *__restrict__ pose, const float *__restrict__ limits,
                                    const float *__restrict__ eps, const float *__restrict__ pad,
                                    const float *__restrict__ clr, const float *__restrict__ dis, int num_objects,
                                    ObjRec *__restrict__ out, long long dil_obj_stride, long long quad_obj_stride) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= num_objects) return;
    const float *P = pose + 16 * o;
    float m[3][3] = {{P[0], P[1], P[2]}, {P[4], P[5], P[6]}, {P[8], P[9], P[10]}};
    float q[3], w;
    float t = __fadd_rn(m[0][0], __fadd_rn(m[1][1], m[2][2]));   // trace(): Eigen's unrolled tree redux c0 + (c1 + c2)
    if (t > 0.0f) {
        t = __fsqrt_rn(__fadd_rn(t, 1.0f));
        w = __fmul_rn(0.5f, t);
        t = __fdiv_rn(0.5f, t);
        q[0] = __fmul_rn(__fsub_rn(m[2][1], m[1][2]), t);
        q[1] = __fmul_rn(__fsub_rn(m[0][2], m[2][0]), t);
        q[2] = __fmul_rn(__fsub_rn(m[1][0], m[0][1]), t);
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = __fsqrt_rn(__fadd_rn(__fsub_rn(__fsub_rn(m[i][i], m[j][j]), m[k][k]), 1.0f));
        q[i] = __fmul_rn(0.5f, t);
        t = __fdiv_rn(0.5f, t);
        w = __fmul_rn(__fsub_rn(m[k][j], m[j][k]), t);
        q[j] = __fmul_rn(__fadd_rn(m[j][i], m[i][j]), t);
        q[k] = __fmul_rn(__fadd_rn(m[k][i], m[i][k]), t);
    }
    ObjRec r;
    r.qw = w; r.qx = q[0]; r.qy = q[1]; r.qz = q[2];
    r.tx = P[3]; r.ty = P[7]; r.tz = P[11];
    // toRotationMatrix with the multiply-adds fused exactly where nvcc fuses them in the reference kernel
    // (oracle/sdf_ref: SASS of the reference source built for sm_100a; tests/test_gpu_ref_operator.py)
    const float tx = __fmul_rn(2.0f, q[0]), ty = __fmul_rn(2.0f, q[1]), tz = __fmul_rn(2.0f, q[2]);
    const float twx = __fmul_rn(tx, w), twz = __fmul_rn(tz, w), txz = __fmul_rn(tz, q[0]);
    const float tyy = __fmul_rn(ty, q[1]), tzz = __fmul_rn(tz, q[2]);
    r.r[0] = __fsub_rn(1.0f, __fadd_rn(tyy, tzz)); r.r[1] = __fmaf_rn(ty, q[0], -twz); r.r[2] = __fmaf_rn(ty, w, txz);
    r.r[3] = __fmaf_rn(ty, q[0], twz); r.r[4] = __fsub_rn(1.0f, __fmaf_rn(tx, q[0], tzz)); r.r[5] = __fmaf_rn(tz, q[1], -twx);
    r.r[6] = __fmaf_rn(-ty, w, txz); r.r[7] = __fmaf_rn(tz, q[1], twx); r.r[8] = __fsub_rn(1.0f, __fmaf_rn(tx, q[0], tyy));
    const float *lim = limits + 10 * o;
    r.minx = lim[0]; r.miny = lim[1]; r.minz = lim[2];
    r.ex = __fsub_rn(lim[3], lim[0]); r.ey = __fsub_rn(lim[4], lim[1]); r.ez = __fsub_rn(lim[5], lim[2]);
    r.d0 = (int)lim[6]; r.d1 = (int)lim[7]; r.d2 = (int)lim[8];
    r.fd0 = (float)r.d0; r.fd1 = (float)r.d1; r.fd2 = (float)r.d2;
    r.delta = lim[9];
    r.eps = eps[o]; r.pad = pad[o]; r.clr = clr[o]; r.dis = dis[o];
    r.inv2eps = __fdiv_rn(1.0f, __fmul_rn(2.0f, r.eps));
    r.inveps = __fdiv_rn(1.0f, r.eps);
    const float sx = r.ex / r.fd0, sy = r.ey / r.fd1, sz = r.ez / r.fd2;
    r.lox = r.minx - 0.5f * sx; r.hix = r.minx + (r.fd0 - 0.5f) * sx;
    r.loy = r.miny - 0.5f * sy; r.hiy = r.miny + (r.fd1 - 0.5f) * sy;
    r.loz = r.minz - 0.5f * sz; r.hiz = r.minz + (r.fd2 - 0.5f) * sz;
    // An out-of-bounds sample reads 1.0 (kernel.cu:47-48): it contributes nothing iff eps < 1 and
    // clearance <= 1.  Otherwise the cull must never reject.
    r.cull_pad = (r.eps < 1.0f && r.clr <= 1.0f) ? (1e-3f + fmaxf(sx, fmaxf(sy, sz))) : 1e30f;
    r.isx = 1.0f / sx; r.isy = 1.0f / sy; r.isz = 1.0f / sz;
    r.alox = r.aloy = r.aloz = -1e30f; r.ahix = r.ahiy = r.ahiz = 1e30f;   // (no active box yet)
    r.dil_off = (int)(dil_obj_stride * o);
    // approximate world -> grid map for the cull / classification tests (their margins absorb its rounding)
    r.ga[0] = r.r[0] * r.isx; r.ga[1] = r.r[1] * r.isx; r.ga[2] = r.r[2] * r.isx; r.ga[3] = (r.tx - r.minx) * r.isx;
    r.ga[4] = r.r[3] * r.isy; r.ga[5] = r.r[4] * r.isy; r.ga[6] = r.r[5] * r.isy; r.ga[7] = (r.ty - r.miny) * r.isy;
    r.ga[8] = r.r[6] * r.isz; r.ga[9] = r.r[7] * r.isz; r.ga[10] = r.r[8] * r.isz; r.ga[11] = (r.tz - r.minz) * r.isz;
    {   // world-frame sphere around the region in which a sample can be in bounds, padded like the box test
        const float cx = 0.5f * (r.lox + r.hix) - r.tx, cy = 0.5f * (r.loy + r.hiy) - r.ty, cz = 0.5f * (r.loz + r.hiz) - r.tz;
        r.wsx = r.r[0] * cx + r.r[3] * cy + r.r[6] * cz;   // R^T (c - t)
        r.wsy = r.r[1] * cx + r.r[4] * cy + r.r[7] * cz;
        r.wsz = r.r[2] * cx + r.r[5] * cy + r.r[8] * cz;
        const float hx = 0.5f * (r.hix - r.lox), hy = 0.5f * (r.hiy - r.loy), hz = 0.5f * (r.hiz - r.loz);
        // 1.001 + 1e-3: fp32 rounding of the (nearly orthonormal) rotation and of the centre
        r.wsr = (r.cull_pad < 1e29f) ? (sqrtf(hx * hx + hy * hy + hz * hz) * 1.001f + r.cull_pad + 1e-3f) : -1.0f;
    }
    r.quad_offset = quad_obj_stride * o;
    r.grid_offset = (long long)o * r.d0 * r.d1 * r.d2;
    out[o] = r;
}

// ----------------------------------------------------------------------------------------------------
// lower-bound grid (DilDesc): 2^3-voxel brick minima, then the minimum over each brick's 3^3 neighbourhood.
// Built once per omgb_scene_set_sdf.
// ----------------------------------------------------------------------------------------------------
__global__ void brick_min_kernel(const float *__restrict__ grids, int num_objects, int d0, int d1, int d2,
                                 float *__restrict__ out, int b0, int b1, int b2) {
    const long long total = (long long)num_objects * b0 * b1 * b2;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int bz = (int)(t % b2);
        const int by = (int)((t / b2) % b1);
        const int bx = (int)((t / ((long long)b2 * b1)) % b0);
        const int o = (int)(t / ((long long)b2 * b1 * b0));
        const float *g = grids + (size_t)o * d0 * d1 * d2;
        float m = 3.0e38f;
        for (int x = bx * 2; x < min(bx * 2 + 2, d0); ++x)
            for (int y = by * 2; y < min(by * 2 + 2, d1); ++y)
                for (int z = bz * 2; z < min(bz * 2 + 2, d2); ++z) m = fminf(m, g[((size_t)x * d1 + y) * d2 + z]);
        out[t] = m;
    }
}

__global__ void brick_dilate_kernel(const float *__restrict__ in, int num_objects, int b0, int b1, int b2,
                                    float *__restrict__ out) {
    const long long total = (long long)num_objects * b0 * b1 * b2;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int bz = (int)(t % b2);
        const int by = (int)((t / b2) % b1);
        const int bx = (int)((t / ((long long)b2 * b1)) % b0);
        const int o = (int)(t / ((long long)b2 * b1 * b0));
        const float *g = in + (size_t)o * b0 * b1 * b2;
        float m = 3.0e38f;
        for (int x = max(bx - 1, 0); x <= min(bx + 1, b0 - 1); ++x)
            for (int y = max(by - 1, 0); y <= min(by + 1, b1 - 1); ++y)
                for (int z = max(bz - 1, 0); z <= min(bz + 1, b2 - 1); ++z) m = fminf(m, g[((size_t)x * b1 + y) * b2 + z]);
        out[t] = m;
    }
}

// Bricked quad copy of the packed grids (QuadDesc, sdf_device.cuh): one thread per output float4, consecutive threads
// write consecutive cells (x fastest inside an 8^3 brick).  Cells whose +y / +z neighbour does not exist are never
// sampled (the in-bounds test comes first); they hold the pad value 1.0.
__global__ void __launch_bounds__(256) quad_pack_kernel(const float *__restrict__ grids, int num_objects, int d0, int d1,
                                                        int d2, int nb0, int nb1, int nb2, float4 *__restrict__ out) {
    const long long per = (long long)nb0 * nb1 * nb2 * 512;
    const long long total = per * num_objects;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(t / per);
        const long long r = t - (long long)o * per;
        const int in = (int)(r & 511);
        const long long brick = r >> 9;
        const int bz = (int)(brick % nb2), by = (int)((brick / nb2) % nb1), bx = (int)(brick / ((long long)nb2 * nb1));
        const int x = bx * 8 + (in & 7), z = bz * 8 + ((in >> 3) & 7), y = by * 8 + (in >> 6);
        float4 v = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        if (x < d0 && y < d1 && z < d2) {
            const float *g = grids + (size_t)o * d0 * d1 * d2 + ((size_t)x * d1 + y) * d2 + z;
            v.x = g[0];
            if (z + 1 < d2) v.y = g[1];
            if (y + 1 < d1) {
                v.z = g[d2];
                if (z + 1 < d2) v.w = g[d2 + 1];
            }
        }
        out[t] = v;
    }
}

// Active box per object: AABB (object frame) of the bricks whose lower bound allows value <= eps or < clearance.
__global__ void active_bounds_init_kernel(int *__restrict__ bounds, int num_objects) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= num_objects) return;
    bounds[6 * o + 0] = bounds[6 * o + 1] = bounds[6 * o + 2] = 0x7fffffff;
    bounds[6 * o + 3] = bounds[6 * o + 4] = bounds[6 * o + 5] = -1;
}

__global__ void __launch_bounds__(256) active_bounds_kernel(const float *__restrict__ dil, const ObjRec *__restrict__ objs,
                                                            int num_objects, int b0, int b1, int b2,
                                                            int *__restrict__ bounds) {
    // grid = (blocks per object, objects): every block reduces its bricks in registers / shared memory and issues at
    // most six global atomics (the naive per-brick atomics serialised on six addresses: 8.7 ms for 10 x 64^3 bricks)
    __shared__ int sb[6];
    const int o = blockIdx.y;
    if (threadIdx.x < 3) sb[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) sb[threadIdx.x] = -1;
    __syncthreads();
    const long long per = (long long)b0 * b1 * b2;
    const float eps = objs[o].eps, clr = objs[o].clr;
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-1, -1, -1};
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < per; r += (long long)gridDim.x * blockDim.x) {
        const float v = dil[(long long)o * per + r];
        const float slack = 1e-4f + 1e-5f * fabsf(v);
        if ((v > eps + slack) && (v > clr + slack)) continue;   // provably inactive brick
        const int bz = (int)(r % b2), by = (int)((r / b2) % b1), bx = (int)(r / ((long long)b2 * b1));
        lo[0] = min(lo[0], bx); lo[1] = min(lo[1], by); lo[2] = min(lo[2], bz);
        hi[0] = max(hi[0], bx); hi[1] = max(hi[1], by); hi[2] = max(hi[2], bz);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
        hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
    }
    if ((threadIdx.x & 31) == 0 && hi[0] >= 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&sb[k], lo[k]); atomicMax(&sb[3 + k], hi[k]); }
    }
    __syncthreads();
    if (threadIdx.x < 3) { if (sb[threadIdx.x] != 0x7fffffff) atomicMin(bounds + 6 * o + threadIdx.x, sb[threadIdx.x]); }
    else if (threadIdx.x < 6) { if (sb[threadIdx.x] >= 0) atomicMax(bounds + 6 * o + threadIdx.x, sb[threadIdx.x]); }
}

__global__ void active_bounds_final_kernel(const int *__restrict__ bounds, ObjRec *__restrict__ objs, int num_objects) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= num_objects) return;
    ObjRec &r = objs[o];
    const float sx = r.ex / r.fd0, sy = r.ey / r.fd1, sz = r.ez / r.fd2;
    // a sample whose grid coordinate g has floor(g) in brick b lies in [2b, 2b+2) voxel units
    r.alox = r.minx + (2.0f * bounds[6 * o + 0] - 0.01f) * sx; r.ahix = r.minx + (2.0f * bounds[6 * o + 3] + 2.01f) * sx;
    r.aloy = r.miny + (2.0f * bounds[6 * o + 1] - 0.01f) * sy; r.ahiy = r.miny + (2.0f * bounds[6 * o + 4] + 2.01f) * sy;
    r.aloz = r.minz + (2.0f * bounds[6 * o + 2] - 0.01f) * sz; r.ahiz = r.minz + (2.0f * bounds[6 * o + 5] + 2.01f) * sz;
    if (bounds[6 * o + 3] < 0) { r.alox = r.aloy = r.aloz = 1e30f; r.ahix = r.ahiy = r.ahiz = -1e30f; }
}

// Longest-processing-time-first CTA order for the next launch: 64-bucket counting sort of the per-trajectory
// CTA clocks of the previous launch, descending.  A scheduling hint only (trajectories are independent).
__global__ void lpt_order_kernel(const int *__restrict__ cost, int *__restrict__ order, int batch) {
    __shared__ int hist[64], base[64], cmax;
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) cmax = 1;
    __syncthreads();
    int m = 1;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) m = max(m, cost[b]);
    atomicMax(&cmax, m);
    __syncthreads();
    const float scale = 63.999f / (float)cmax;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) atomicAdd(&hist[63 - (int)(cost[b] * scale)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < 64; ++k) { base[k] = acc; acc += hist[k]; }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < batch; b += blockDim.x) order[atomicAdd(&base[63 - (int)(cost[b] * scale)], 1)] = b;
}

// ----------------------------------------------------------------------------------------------------
// raw operator: drop-in for omg_cuda.sdf_loss_forward (one thread per point, objects in ascending order)
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sdf_loss_kernel(const ObjRec *__restrict__ objs, int num_objects,
                                                       const float *__restrict__ grids,
                                                       const float *__restrict__ points, int num_points,
                                                       float *__restrict__ potentials,
                                                       float *__restrict__ potential_grads,
                                                       float *__restrict__ collides) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ObjRec *s_objs = reinterpret_cast<ObjRec *>(smem_raw);
    {
        const int words = (int)(sizeof(ObjRec) / 4) * num_objects;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(objs);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_objs);
        for (int k = threadIdx.x; k < words; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < num_points; n += gridDim.x * blockDim.x) {
        const float x = points[3 * n], y = points[3 * n + 1], z = points[3 * n + 2];
        float pot = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f, col = 0.0f;
        for (int o = 0; o < num_objects; ++o) {
            const ObjRec &ob = s_objs[o];
            if (ob.dis > 0.0f) continue;   // kernel.cu:115
            float po, ax, ay, az, co;
            pair_full(ob, grids, QuadDesc{nullptr, 0, 0, 0}, x, y, z, po, ax, ay, az, co);
            pot = __fadd_rn(pot, po);
            gx = __fadd_rn(gx, ax); gy = __fadd_rn(gy, ay); gz = __fadd_rn(gz, az);
            col = __fadd_rn(col, co);
        }
        potentials[n] = pot;
        potential_grads[3 * n] = gx; potential_grads[3 * n + 1] = gy; potential_grads[3 * n + 2] = gz;
        collides[n] = col;
    }
}

// ----------------------------------------------------------------------------------------------------
// Cost.batch_obstacle_cost (omg/cost.py:192-286): FK + operator (+ arc-length weighting) for M configs
// ----------------------------------------------------------------------------------------------------
constexpr int BOC_CFG = 8;   // configurations per CTA

__global__ void __launch_bounds__(256) batch_obstacle_cost_kernel(
    const ObjRec *__restrict__ objs, int num_objects, const float *__restrict__ grids,
    const RobotConst *__restrict__ rc, const double *__restrict__ joints, int num_configs, int arc_length,
    const double *__restrict__ start, float inv_dt, int finger_soft, float *__restrict__ potentials,
    float *__restrict__ grads, float *__restrict__ collides) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ObjRec *s_objs = reinterpret_cast<ObjRec *>(smem_raw);
    double *s_frames = reinterpret_cast<double *>(smem_raw + align_up(sizeof(ObjRec) * num_objects, 16));
    // frames of configs m0-1 (halo / start), m0 .. m0+BOC_CFG-1
    const int m0 = blockIdx.x * BOC_CFG;
    {
        const int words = (int)(sizeof(ObjRec) / 4) * num_objects;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(objs);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_objs);
        for (int k = threadIdx.x; k < words; k += blockDim.x) dst[k] = src[k];
    }
    if (threadIdx.x < BOC_CFG + 2) {
        // slot 0: config m0-1 (halo), slots 1..BOC_CFG: configs m0.., slot BOC_CFG+1: FK(start)
        const int m = m0 - 1 + (int)threadIdx.x;
        const double *q = nullptr;
        if (threadIdx.x == BOC_CFG + 1) {
            if (arc_length > 0) q = start;
        } else if (threadIdx.x == 0) {
            if (arc_length > 0 && m >= 0) q = joints + (size_t)m * ND;
        } else if (m < num_configs) {
            q = joints + (size_t)m * ND;
        }
        if (q) {
            double ql[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) ql[d] = q[d];
            panda_fk(rc, ql, s_frames + (size_t)threadIdx.x * NL * 12);
        }
    }
    __syncthreads();
    const int P = rc->p;
    const int per_cfg = NL * P;
    for (int k = threadIdx.x; k < BOC_CFG * per_cfg; k += blockDim.x) {
        const int lc = k / per_cfg, r = k - lc * per_cfg;
        const int m = m0 + lc;
        if (m >= num_configs) break;
        const int j = r / P, p = r - j * P;
        const double *F = s_frames + ((size_t)(lc + 1) * NL + j) * 12;
        double X, Y, Z;
        xform(F, rc->pts[j][p][0], rc->pts[j][p][1], rc->pts[j][p][2], X, Y, Z);
        const float x = (float)X, y = (float)Y, z = (float)Z;
        float pot = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f, col = 0.0f;
        for (int o = 0; o < num_objects; ++o) {
            const ObjRec &ob = s_objs[o];
            if (ob.dis > 0.0f) continue;
            float po, ax, ay, az, co;
            pair_full(ob, grids, QuadDesc{nullptr, 0, 0, 0}, x, y, z, po, ax, ay, az, co);
            pot = __fadd_rn(pot, po);
            gx = __fadd_rn(gx, ax); gy = __fadd_rn(gy, ay); gz = __fadd_rn(gz, az);
            col = __fadd_rn(col, co);
        }
        if (finger_soft && j >= 8) {
            pot = __fmul_rn(pot, 0.1f); gx = __fmul_rn(gx, 0.1f); gy = __fmul_rn(gy, 0.1f);
            gz = __fmul_rn(gz, 0.1f); col = 0.0f;
        }
        if (arc_length > 0) {   // potential x |workspace velocity| (cost.py:235-275, config.py:162-187), fp32
            // previous configuration of the same group, or FK(start) for the first of a group
            const bool first_in_group = (m % arc_length) == 0;
            const double *Fp = s_frames + ((size_t)(first_in_group ? BOC_CFG + 1 : lc) * NL + j) * 12;
            double Xp, Yp, Zp;
            xform(Fp, rc->pts[j][p][0], rc->pts[j][p][1], rc->pts[j][p][2], Xp, Yp, Zp);
            const float vx = (x - (float)Xp) * inv_dt, vy = (y - (float)Yp) * inv_dt, vz = (z - (float)Zp) * inv_dt;
            pot *= sqrtf(vx * vx + vy * vy + vz * vz);
        }
        const size_t idx = (size_t)m * per_cfg + r;
        potentials[idx] = pot;
        collides[idx] = col;
        if (grads) { grads[3 * idx] = gx; grads[3 * idx + 1] = gy; grads[3 * idx + 2] = gz; }
    }
}

}  // namespace omgb

using namespace omgb;

// ----------------------------------------------------------------------------------------------------
// scene
// ----------------------------------------------------------------------------------------------------
constexpr int ORDER_SLOTS = 8;
constexpr int PIPE_CHUNKS = 4;
struct OrderSlot {
    int *d_order = nullptr, *d_cost = nullptr;
    int cap = 0, batch = -1;
    const void *key = nullptr;
    bool valid = false;
    unsigned age = 0;
};
constexpr unsigned LPT_REFRESH = 4;

struct omgb_scene {
    int device = 0;
    RobotConst *d_robot = nullptr;
    RobotParams rp;
    bool robot_set = false;
    int p = 0;
    const float *d_grids = nullptr;
    float *d_limits = nullptr;
    int num_objects = 0, gx = 0, gy = 0, gz = 0;
    bool sdf_set = false;
    float *d_objparams = nullptr;   // pose[16 O] eps[O] pad[O] clr[O] dis[O]
    ObjRec *d_objs = nullptr;
    bool objs_set = false;
    double *d_Ainv = nullptr, *d_proj = nullptr;
    int n = 0, c = 0;
    bool metric_set = false;
    int metric_kind = 0;          // closed form of Ainv recognised by omgb_scene_set_metric (0 = none, dense product)
    double metric_scale = 0.0;
    // staging for the host-buffer entry point
    float *d_dil = nullptr;
    DilDesc dil;
    float4 *d_quad = nullptr;     // bricked quad copy of the grids (omgb_scene_set_sdf_layout), or null
    QuadDesc quad;
    int *d_bounds = nullptr;
    // longest-first CTA scheduling state (hint only), one slot per (xi buffer, batch) seen recently
    OrderSlot order[ORDER_SLOTS];
    int order_next = 0;
    int use_lpt = 1;
    // host-buffer entry point: 0 auto (zero-copy when every buffer is mapped pinned memory, else pipelined staging),
    // 1 staged in one piece, 2 staged + pipelined over chunks, 3 zero-copy required
    int host_mode = 0;
    // persistent plan kernel state: schedules [iters][3], per-trajectory progress, the item counter
    double *d_plan_sched = nullptr;
    int *d_plan_progress = nullptr;
    unsigned *d_plan_counter = nullptr;
    int plan_sched_cap = 0, plan_progress_cap = 0;
    int num_sms = 148;
    cudaStream_t pipe_stream[PIPE_CHUNKS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pipe_done[PIPE_CHUNKS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pipe_begin = nullptr;
    long long *d_prof = nullptr;   // diagnostic: per-CTA phase clocks (omgb_scene_set_profile)
    double *d_stage = nullptr;
    size_t stage_bytes = 0;
    int smem_optin = 0;
    // per-point potentials of the top-k path: [trajectories][n*10][LPI] fp32, grown on demand (launch_step)
    float *d_pot = nullptr;
    size_t pot_floats = 0;
    // One engine is one stream's worth of mutable state (CTA order slots, plan counters, the potential scratch).
    // Every launch records `last_done` on its stream; a launch arriving on ANOTHER stream waits for it first, so
    // driving one engine from several streams is serialised instead of racing (use one engine per stream to overlap).
    cudaEvent_t last_done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
};

// Orders a launch on `st` behind the previous launch of this scene when that ran on a different stream.
static int stream_enter(omgb_scene *s, cudaStream_t st) {
    if (s->has_last && s->last_stream != st) {
        cudaError_t e = cudaStreamWaitEvent(st, s->last_done, 0);
        if (e != cudaSuccess) return host_fail(OMGB_ERR_CUDA, std::string("cudaStreamWaitEvent: ") + cudaGetErrorString(e));
    }
    return OMGB_OK;
}
static int stream_leave(omgb_scene *s, cudaStream_t st) {
    if (!s->last_done) {
        cudaError_t e = cudaEventCreateWithFlags(&s->last_done, cudaEventDisableTiming);
        if (e != cudaSuccess) return host_fail(OMGB_ERR_CUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(e));
    }
    cudaError_t e = cudaEventRecord(s->last_done, st);
    if (e != cudaSuccess) return host_fail(OMGB_ERR_CUDA, std::string("cudaEventRecord: ") + cudaGetErrorString(e));
    s->last_stream = st;
    s->has_last = true;
    return OMGB_OK;
}

extern "C" int omgb_version(void) { return OMGB_VERSION; }
extern "C" const char *omgb_last_error(void) { return g_err.c_str(); }

extern "C" int omgb_scene_create(omgb_scene_t **out, int device) {
    if (!out) return fail(OMGB_ERR_INVALID, "omgb_scene_create: out is null");
    OMGB_CUDA(cudaSetDevice(device));
    omgb_scene *s = new omgb_scene();
    s->device = device;
    memset(&s->dil, 0, sizeof(s->dil));
    memset(&s->quad, 0, sizeof(s->quad));
    cudaError_t e = cudaMalloc(&s->d_robot, sizeof(RobotConst));
    if (e != cudaSuccess) { delete s; return fail(OMGB_ERR_CUDA, cudaGetErrorString(e)); }
    cudaDeviceGetAttribute(&s->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, device);
    *out = s;
    return OMGB_OK;
}

extern "C" int omgb_scene_destroy(omgb_scene_t *s) {
    if (!s) return OMGB_OK;
    cudaSetDevice(s->device);
    cudaFree(s->d_robot); cudaFree(s->d_limits); cudaFree(s->d_objparams); cudaFree(s->d_objs);
    cudaFree(s->d_Ainv); cudaFree(s->d_proj); cudaFree(s->d_stage); cudaFree(s->d_dil); cudaFree(s->d_bounds);
    cudaFree(s->d_quad);
    for (int k = 0; k < ORDER_SLOTS; ++k) { cudaFree(s->order[k].d_order); cudaFree(s->order[k].d_cost); }
    cudaFree(s->d_plan_sched); cudaFree(s->d_plan_progress); cudaFree(s->d_plan_counter); cudaFree(s->d_pot);
    if (s->last_done) cudaEventDestroy(s->last_done);
    for (int k = 0; k < PIPE_CHUNKS; ++k) {
        if (s->pipe_stream[k]) cudaStreamDestroy(s->pipe_stream[k]);
        if (s->pipe_done[k]) cudaEventDestroy(s->pipe_done[k]);
    }
    if (s->pipe_begin) cudaEventDestroy(s->pipe_begin);
    delete s;
    return OMGB_OK;
}

static void mat34(const double *m44, double *out12) {
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) out12[3 * r + c] = m44[4 * r + c];
        out12[9 + r] = m44[4 * r + 3];
    }
}

extern "C" int omgb_scene_set_robot(omgb_scene_t *s, const double *pose_0, const double *tip2joint,
                                    const double *joint_axis, const double *joint_origin, int use_true_origin,
                                    const double *center_offset, const double *body_points, int p,
                                    const double *lower, const double *upper) {
    if (!s || !pose_0 || !tip2joint || !joint_axis || !center_offset || !body_points || !lower || !upper)
        return fail(OMGB_ERR_INVALID, "omgb_scene_set_robot: null argument");
    if (p < 1 || p > OMGB_MAX_BODY_POINTS) return fail(OMGB_ERR_INVALID, "points_per_link out of range");
    if (use_true_origin && !joint_origin) return fail(OMGB_ERR_INVALID, "joint_origin required");
    OMGB_CUDA(cudaSetDevice(s->device));
    std::vector<RobotConst> hv(1);
    RobotConst &h = hv[0];
    memset(&h, 0, sizeof(h));
    for (int i = 0; i < 10; ++i) {
        mat34(pose_0 + 16 * i, h.P0[i]);
        mat34(center_offset + 16 * i, h.CO[i]);
        const double *t2j = tip2joint + 16 * i;
        const double *ax = joint_axis + 3 * i;
        const double *og = use_true_origin ? joint_origin + 3 * i : ax;   // robot_pykdl.py:104 aliasing
        // joint axis / "origin" in the joint's link frame (joint_pose = T * tip2joint, robot_pykdl.py:193-202),
        // then re-expressed in the body-point frame F = T * center_offset the kernel stores:
        //   axis_world = F.R * (CO_R^T a),  origin_world = F.R * (CO_R^T (o - CO_t)) + F.t
        double ja[3], jo[3];
        for (int r = 0; r < 3; ++r) {
            ja[r] = t2j[4 * r] * ax[0] + t2j[4 * r + 1] * ax[1] + t2j[4 * r + 2] * ax[2];
            jo[r] = t2j[4 * r] * og[0] + t2j[4 * r + 1] * og[1] + t2j[4 * r + 2] * og[2] + t2j[4 * r + 3];
        }
        for (int r = 0; r < 3; ++r) {
            const double *co = h.CO[i];
            h.jab[i][r] = co[r] * ja[0] + co[3 + r] * ja[1] + co[6 + r] * ja[2];
            h.job[i][r] = co[r] * (jo[0] - co[9]) + co[3 + r] * (jo[1] - co[10]) + co[6 + r] * (jo[2] - co[11]);
        }
        // bounding sphere of the link's body points (centre = midpoint of the AABB)
        double lo[3] = {1e30, 1e30, 1e30}, hi[3] = {-1e30, -1e30, -1e30};
        for (int k = 0; k < p; ++k)
            for (int r = 0; r < 3; ++r) {
                const double v = body_points[((size_t)i * p + k) * 3 + r];
                h.pts[i][k][r] = v;
                lo[r] = v < lo[r] ? v : lo[r];
                hi[r] = v > hi[r] ? v : hi[r];
            }
        double cx[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])}, rad = 0;
        for (int k = 0; k < p; ++k) {
            double d2 = 0;
            for (int r = 0; r < 3; ++r) { const double d = h.pts[i][k][r] - cx[r]; d2 += d * d; }
            rad = d2 > rad ? d2 : rad;
        }
        for (int r = 0; r < 3; ++r) h.sph[i][r] = (float)cx[r];
        h.sph[i][3] = (float)(sqrt(rad) * 1.0001 + 1e-6);
    }
    for (int d = 0; d < ND; ++d) { h.lower[d] = lower[d]; h.upper[d] = upper[d]; }
    h.p = p;
    OMGB_CUDA(cudaMemcpy(s->d_robot, &h, sizeof(h), cudaMemcpyHostToDevice));
    memcpy(s->rp.P0, h.P0, sizeof(h.P0));
    memcpy(s->rp.CO, h.CO, sizeof(h.CO));
    memcpy(s->rp.lower, h.lower, sizeof(h.lower));
    memcpy(s->rp.upper, h.upper, sizeof(h.upper));
    memcpy(s->rp.sph, h.sph, sizeof(h.sph));
    s->p = p;
    s->robot_set = true;
    return OMGB_OK;
}

extern "C" unsigned long long omgb_launch_count(void) { return g_launches.load(); }

extern "C" int omgb_scene_set_options(omgb_scene_t *s, int use_lower_bound, int use_longest_first) {
    if (!s) return fail(OMGB_ERR_INVALID, "omgb_scene_set_options: null scene");
    if (use_lower_bound >= 0) s->dil.enabled = (use_lower_bound != 0 && s->d_dil != nullptr) ? 1 : 0;
    if (use_longest_first >= 0) {
        s->use_lpt = use_longest_first != 0;
        for (int k = 0; k < ORDER_SLOTS; ++k) s->order[k].valid = false;
    }
    return OMGB_OK;
}

extern "C" int omgb_scene_set_host_mode(omgb_scene_t *s, int mode) {
    if (!s || mode < 0 || mode > 3) return fail(OMGB_ERR_INVALID, "omgb_scene_set_host_mode: bad argument");
    s->host_mode = mode;
    return OMGB_OK;
}

extern "C" int omgb_scene_set_profile(omgb_scene_t *s, long long *d_phase_clocks) {
    if (!s) return fail(OMGB_ERR_INVALID, "omgb_scene_set_profile: null scene");
    s->d_prof = d_phase_clocks;
    return OMGB_OK;
}

extern "C" int omgb_scene_set_sdf(omgb_scene_t *s, const float *d_sdf_grids, const float *h_sdf_limits,
                                  int num_objects, int dx, int dy, int dz, void *stream) {
    if (!s || !d_sdf_grids || !h_sdf_limits) return fail(OMGB_ERR_INVALID, "omgb_scene_set_sdf: null argument");
    if (num_objects < 1 || num_objects > OMGB_MAX_OBJECTS)
        return fail(OMGB_ERR_INVALID, "num_objects must be in [1, OMGB_MAX_OBJECTS]");
    for (int o = 0; o < num_objects; ++o) {
        const float *l = h_sdf_limits + 10 * o;
        if ((int)l[6] != dx || (int)l[7] != dy || (int)l[8] != dz)
            return fail(OMGB_ERR_INVALID, "sdf_limits dims disagree with the packed grid shape");
    }
    OMGB_CUDA(cudaSetDevice(s->device));
    if (num_objects != s->num_objects) {
        cudaFree(s->d_limits); cudaFree(s->d_objparams); cudaFree(s->d_objs);
        s->d_limits = nullptr; s->d_objparams = nullptr; s->d_objs = nullptr;
        OMGB_CUDA(cudaMalloc(&s->d_limits, sizeof(float) * 10 * num_objects));
        OMGB_CUDA(cudaMalloc(&s->d_objparams, sizeof(float) * 20 * num_objects));
        OMGB_CUDA(cudaMalloc(&s->d_objs, sizeof(ObjRec) * num_objects));
    }
    OMGB_CUDA(cudaMemcpy(s->d_limits, h_sdf_limits, sizeof(float) * 10 * num_objects, cudaMemcpyHostToDevice));
    s->d_grids = d_sdf_grids;
    s->num_objects = num_objects; s->gx = dx; s->gy = dy; s->gz = dz;
    {   // lower-bound grid (see DilDesc)
        DilDesc &dd = s->dil;
        memset(&dd, 0, sizeof(dd));
        dd.bx = (dx + 1) / 2; dd.by = (dy + 1) / 2; dd.bz = (dz + 1) / 2;
        dd.obj_stride = (long long)dd.bx * dd.by * dd.bz;
        const long long total = dd.obj_stride * num_objects;
        if (total > 0x7fffffffLL) return fail(OMGB_ERR_UNSUPPORTED, "packed SDFs too large for the lower-bound grid index");
        cudaFree(s->d_dil);
        s->d_dil = nullptr;
        float *tmp = nullptr;
        OMGB_CUDA(cudaMalloc(&s->d_dil, sizeof(float) * (size_t)total));
        OMGB_CUDA(cudaMalloc(&tmp, sizeof(float) * (size_t)total));
        const int blocks = (int)((total + 255) / 256 > 148 * 64 ? 148 * 64 : (total + 255) / 256);
        // on the caller's stream: the grid was produced there (omgb_sdf_pack, torch ops); only that stream is waited for
        cudaStream_t st = (cudaStream_t)stream;
        brick_min_kernel<<<blocks, 256, 0, st>>>(d_sdf_grids, num_objects, dx, dy, dz, tmp, dd.bx, dd.by, dd.bz);
        brick_dilate_kernel<<<blocks, 256, 0, st>>>(tmp, num_objects, dd.bx, dd.by, dd.bz, s->d_dil);
        cudaError_t e1 = cudaGetLastError(), e2 = cudaStreamSynchronize(st);
        cudaFree(tmp);
        if (e1 != cudaSuccess || e2 != cudaSuccess)
            return fail(OMGB_ERR_CUDA, std::string("lower-bound grid build: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        dd.data = s->d_dil;
        const char *env = getenv("OMGB_NO_LOWER_BOUND");
        dd.enabled = (env && atoi(env)) ? 0 : 1;
    }
    s->sdf_set = true;
    s->objs_set = false;
    // a quad copy of the previous grids is stale; OMGB_SDF_LAYOUT=1 builds the new one right away (A/B runs)
    cudaFree(s->d_quad);
    s->d_quad = nullptr;
    memset(&s->quad, 0, sizeof(s->quad));
    // Default: grids that cannot stay L2-resident (>= 256 MB) also get the bricked quad copy when 4x their size is
    // at most 16 GB -- measured on the config-4 scene (1.34 GB): 0.2489 vs 0.2555 ms per step, whole plan +6 %;
    // neutral on config 2 (84 MB, L2-resident).  OMGB_SDF_LAYOUT=0 / 1 forces the reference layout / the copy.
    const char *lay = getenv("OMGB_SDF_LAYOUT");
    const double bytes = 4.0 * (double)num_objects * dx * dy * dz;
    const bool want = lay ? (atoi(lay) == 1) : (bytes >= 256e6 && 4.0 * bytes <= 16e9);
    if (want) return omgb_scene_set_sdf_layout(s, 1, stream);
    return OMGB_OK;
}

extern "C" int omgb_scene_set_sdf_layout(omgb_scene_t *s, int layout, void *stream) {
    if (!s || (layout != 0 && layout != 1)) return fail(OMGB_ERR_INVALID, "omgb_scene_set_sdf_layout: layout is 0 or 1");
    if (!s->sdf_set) return fail(OMGB_ERR_STATE, "omgb_scene_set_sdf_layout: call omgb_scene_set_sdf first");
    OMGB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (s->has_last) OMGB_CUDA(cudaEventSynchronize(s->last_done));   // (launches still reading the old copy)
    cudaFree(s->d_quad);
    s->d_quad = nullptr;
    memset(&s->quad, 0, sizeof(s->quad));
    s->objs_set = false;   // the object records carry the per-object offset
    if (layout == 0) return OMGB_OK;
    const int nb0 = (s->gx + 7) / 8, nb1 = (s->gy + 7) / 8, nb2 = (s->gz + 7) / 8;
    const long long per = (long long)nb0 * nb1 * nb2 * 512, total = per * s->num_objects;
    OMGB_CUDA(cudaMalloc(&s->d_quad, sizeof(float4) * (size_t)total));
    const long long want = (total + 255) / 256;
    const int blocks = (int)(want > 148LL * 32 ? 148LL * 32 : want);
    quad_pack_kernel<<<blocks, 256, 0, st>>>(s->d_grids, s->num_objects, s->gx, s->gy, s->gz, nb0, nb1, nb2, s->d_quad);
    ++g_launches;
    OMGB_CUDA(cudaGetLastError());
    OMGB_CUDA(cudaStreamSynchronize(st));
    s->quad.data = s->d_quad; s->quad.obj_stride = per; s->quad.nb1 = nb1; s->quad.nb2 = nb2;
    return OMGB_OK;
}

extern "C" int omgb_scene_set_objects(omgb_scene_t *s, const float *pose_inv, const float *eps, const float *pad,
                                      const float *clr, const float *dis, void *stream) {
    if (!s || !pose_inv || !eps || !pad || !clr || !dis)
        return fail(OMGB_ERR_INVALID, "omgb_scene_set_objects: null argument");
    if (!s->sdf_set) return fail(OMGB_ERR_STATE, "omgb_scene_set_objects: call omgb_scene_set_sdf first");
    OMGB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int O = s->num_objects;
    std::vector<float> h(20 * (size_t)O);
    memcpy(h.data(), pose_inv, sizeof(float) * 16 * O);
    memcpy(h.data() + 16 * O, eps, sizeof(float) * O);
    memcpy(h.data() + 17 * O, pad, sizeof(float) * O);
    memcpy(h.data() + 18 * O, clr, sizeof(float) * O);
    memcpy(h.data() + 19 * O, dis, sizeof(float) * O);
    OMGB_CUDA(cudaMemcpyAsync(s->d_objparams, h.data(), sizeof(float) * 20 * O, cudaMemcpyHostToDevice, st));
    OMGB_CUDA(cudaStreamSynchronize(st));   // h goes out of scope
    prep_objects_kernel<<<(O + 63) / 64, 64, 0, st>>>(s->d_objparams, s->d_limits, s->d_objparams + 16 * O,
                                                      s->d_objparams + 17 * O, s->d_objparams + 18 * O,
                                                      s->d_objparams + 19 * O, O, s->d_objs, s->dil.obj_stride,
                                                      s->quad.obj_stride);
    OMGB_CUDA(cudaGetLastError());
    if (s->dil.enabled) {   // active boxes depend on eps / clearance
        if (!s->d_bounds) OMGB_CUDA(cudaMalloc(&s->d_bounds, sizeof(int) * 6 * OMGB_MAX_OBJECTS));
        const long long per = s->dil.obj_stride;
        long long bpo = (per + 256 * 8 - 1) / (256 * 8);   // ~8 bricks per thread
        bpo = bpo < 1 ? 1 : (bpo > 148 * 4 ? 148 * 4 : bpo);
        active_bounds_init_kernel<<<1, 64, 0, st>>>(s->d_bounds, O);
        active_bounds_kernel<<<dim3((unsigned)bpo, (unsigned)O), 256, 0, st>>>(s->d_dil, s->d_objs, O, s->dil.bx, s->dil.by,
                                                                           s->dil.bz, s->d_bounds);
        active_bounds_final_kernel<<<1, 64, 0, st>>>(s->d_bounds, s->d_objs, O);
        OMGB_CUDA(cudaGetLastError());
    }
    s->objs_set = true;
    return OMGB_OK;
}

extern "C" int omgb_scene_set_metric(omgb_scene_t *s, int n, const double *h_Ainv, int c, const double *h_proj) {
    if (!s || !h_Ainv || n < 2) return fail(OMGB_ERR_INVALID, "omgb_scene_set_metric: bad argument");
    if (c < 0 || c > n || (c > 0 && !h_proj)) return fail(OMGB_ERR_INVALID, "omgb_scene_set_metric: bad constraint rows");
    OMGB_CUDA(cudaSetDevice(s->device));
    cudaFree(s->d_Ainv); cudaFree(s->d_proj);
    s->d_Ainv = nullptr; s->d_proj = nullptr;
    OMGB_CUDA(cudaMalloc(&s->d_Ainv, sizeof(double) * n * n));
    OMGB_CUDA(cudaMemcpy(s->d_Ainv, h_Ainv, sizeof(double) * n * n, cudaMemcpyHostToDevice));
    if (c > 0) {
        OMGB_CUDA(cudaMalloc(&s->d_proj, sizeof(double) * n * c));
        OMGB_CUDA(cudaMemcpy(s->d_proj, h_proj, sizeof(double) * n * c, cudaMemcpyHostToDevice));
    }
    s->n = n; s->c = c;
    s->metric_set = true;
    // The CHOMP metric A = K^T K (omg/config.py:208-220) is tridiagonal: Ainv = dt^2 min(i,j) with a free end
    // (goal-set mode) and dt^2 min(i,j)(n+1-max(i,j))/(n+1) with a fixed end (1-based).  When the given matrix is one
    // of these (to 1e-9 of its largest entry) the kernel applies it as two running sums per DOF; any other matrix
    // takes the dense product.  OMGB_DENSE_METRIC=1 forces the dense product.
    s->metric_kind = 0;
    s->metric_scale = 0.0;
    const char *env = getenv("OMGB_DENSE_METRIC");
    if (!(env && atoi(env))) {
        double amax = 0.0;
        for (int k = 0; k < n * n; ++k) amax = fabs(h_Ainv[k]) > amax ? fabs(h_Ainv[k]) : amax;
        for (int kind = 1; kind <= 2 && s->metric_kind == 0; ++kind) {
            const double scale = kind == 1 ? h_Ainv[0] : h_Ainv[0] * (double)(n + 1) / (double)n;
            bool ok = scale > 0.0;
            for (int i = 1; i <= n && ok; ++i)
                for (int j = 1; j <= n && ok; ++j) {
                    const double mn = i < j ? i : j, mx = i < j ? j : i;
                    const double want = kind == 1 ? scale * mn : scale * mn * (n + 1 - mx) / (double)(n + 1);
                    ok = fabs(h_Ainv[(size_t)(i - 1) * n + (j - 1)] - want) <= 1e-9 * amax;
                }
            if (ok) { s->metric_kind = kind; s->metric_scale = scale; }
        }
    }
    return OMGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// raw operator
// ----------------------------------------------------------------------------------------------------
extern "C" size_t omgb_sdf_loss_workspace_bytes(int num_objects) {
    return sizeof(ObjRec) * (size_t)(num_objects > 0 ? num_objects : 0);
}

extern "C" int omgb_sdf_loss(const float *pose_init, const float *sdf_grids, const float *sdf_limits,
                             const float *points, const float *epsilons, const float *padding_scales,
                             const float *clearances, const float *disables, int num_points, int num_objects,
                             int dx, int dy, int dz, float *potentials, float *potential_grads, float *collides,
                             void *workspace, void *stream) {
    if (!pose_init || !sdf_grids || !sdf_limits || !epsilons || !padding_scales || !clearances || !disables ||
        !workspace)
        return fail(OMGB_ERR_INVALID, "omgb_sdf_loss: null argument");
    if (num_points < 0 || num_objects < 1 || dx < 2 || dy < 2 || dz < 2)
        return fail(OMGB_ERR_INVALID, "omgb_sdf_loss: bad sizes");
    if (num_points == 0) return OMGB_OK;
    if (!points || !potentials || !potential_grads || !collides)
        return fail(OMGB_ERR_INVALID, "omgb_sdf_loss: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    ObjRec *recs = reinterpret_cast<ObjRec *>(workspace);
    prep_objects_kernel<<<(num_objects + 63) / 64, 64, 0, st>>>(pose_init, sdf_limits, epsilons, padding_scales,
                                                                clearances, disables, num_objects, recs, 0, 0);
    OMGB_CUDA(cudaGetLastError());
    const size_t smem = sizeof(ObjRec) * (size_t)num_objects;
    if (smem > 48 * 1024)
        OMGB_CUDA(cudaFuncSetAttribute(sdf_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (num_points + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    sdf_loss_kernel<<<blocks, 256, smem, st>>>(recs, num_objects, sdf_grids, points, num_points, potentials,
                                               potential_grads, collides);
    g_launches += 2;
    OMGB_CUDA(cudaGetLastError());
    return OMGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// fused CHOMP iteration
// ----------------------------------------------------------------------------------------------------
static int check_step(const omgb_scene *s, const omgb_step_params_t *prm, int batch, const char *who) {
    if (!s || !prm) return fail(OMGB_ERR_INVALID, std::string(who) + ": null argument");
    if (!s->robot_set || !s->sdf_set || !s->objs_set || !s->metric_set)
        return fail(OMGB_ERR_STATE, std::string(who) + ": scene needs robot, sdf, objects and metric");
    if (batch < 0) return fail(OMGB_ERR_INVALID, std::string(who) + ": negative batch");
    if (prm->n_waypoints != s->n)
        return fail(OMGB_ERR_INVALID, std::string(who) + ": n_waypoints differs from the metric set on the scene");
    const int c = prm->goal_set_proj ? prm->constraint_rows : 0;
    if (c != s->c || (prm->goal_set_proj && c < 1))
        return fail(OMGB_ERR_INVALID, std::string(who) + ": constraint_rows differs from the metric set on the scene");
    if (prm->top_k_collision < 0) return fail(OMGB_ERR_INVALID, std::string(who) + ": negative top_k_collision");
    if (!(prm->time_interval > 0)) return fail(OMGB_ERR_INVALID, std::string(who) + ": time_interval must be > 0");
    return OMGB_OK;
}

// Shared-memory carveout (percent of 228 KB) that lets `ctas` CTAs of `smem` dynamic bytes (+1 KB reserved each)
// be resident; whatever is left of the 256 KB array stays L1.
static int carveout_percent(size_t smem, int ctas) {
    const double need = (double)ctas * (double)(smem + 1024);
    int pct = (int)(need / (228.0 * 1024.0) * 100.0) + 1;
    return pct > 100 ? 100 : pct;
}

template <int LPI, int THREADS, int MINB, bool TOPK>
static int launch_one(const StepArgs &a, size_t smem, cudaStream_t st, const PlanArgs *plan, int num_sms) {
    // function attributes are per (instantiation, device); set again only when the footprint changes
    static size_t cached_smem[2][64] = {{0}};
    int dev = 0;
    cudaGetDevice(&dev);
    const int which = plan ? 1 : 0;
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    if (dev < 0 || dev >= 64 || cached_smem[which][dev] != smem) {
        if (plan) {
            OMGB_CUDA(cudaFuncSetAttribute(chomp_plan_kernel<LPI, THREADS, MINB, TOPK>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            OMGB_CUDA(cudaFuncSetAttribute(chomp_plan_kernel<LPI, THREADS, MINB, TOPK>,
                                           cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent(smem, MINB)));
        } else {
            OMGB_CUDA(cudaFuncSetAttribute(chomp_step_kernel<LPI, THREADS, MINB, TOPK>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            OMGB_CUDA(cudaFuncSetAttribute(chomp_step_kernel<LPI, THREADS, MINB, TOPK>,
                                           cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent(smem, MINB)));
        }
        if (dev >= 0 && dev < 64) cached_smem[which][dev] = smem;
    }
    if (plan) {
        const int slots = MINB * num_sms;   // one persistent CTA per resident slot
        chomp_plan_kernel<LPI, THREADS, MINB, TOPK><<<a.batch < slots ? a.batch : slots, THREADS, smem, st>>>(a, *plan);
    } else {
        chomp_step_kernel<LPI, THREADS, MINB, TOPK><<<a.batch, THREADS, smem, st>>>(a);
    }
    ++g_launches;
    OMGB_CUDA(cudaGetLastError());
    return OMGB_OK;
}

template <int LPI, int THREADS, int MINB>
static int launch_cfg(const StepArgs &a, size_t smem, cudaStream_t st, const PlanArgs *plan, int num_sms) {
    return a.prm.top_k_collision > 0 ? launch_one<LPI, THREADS, MINB, true>(a, smem, st, plan, num_sms)
                                     : launch_one<LPI, THREADS, MINB, false>(a, smem, st, plan, num_sms);
}

static int step_config() {   // OMGB_STEP_CONFIG: 0 = 320 threads x 3 CTAs/SM, 1 = 512 x 2, 2 = 1024 x 1, 3 = 256 x 4; else by footprint
    static int cfg = -2;
    if (cfg == -2) {
        const char *e = getenv("OMGB_STEP_CONFIG");
        cfg = e ? atoi(e) : -1;
    }
    return cfg;
}

// plan == nullptr: one iteration, CTA per trajectory; else the persistent plan kernel.
// pot_first: index of this launch's first trajectory inside the scene's potential scratch (the chunks of the pipelined
// host entry point run concurrently on their own streams, each on its own slice).
static int launch_step(omgb_scene *s, const StepArgs &a_in, cudaStream_t st, const PlanArgs *plan = nullptr,
                       size_t pot_first = 0, bool guard_stream = true) {
    const StepArgs &a0 = a_in;
    const int lpi = s->p <= 16 ? 16 : 32;
    const int n = a0.prm.n_waypoints, c = a0.prm.constraint_rows;
    const bool topk = a0.prm.top_k_collision > 0, fing = a0.prm.consider_finger != 0;
    if (a0.batch == 0) return OMGB_OK;
    // CTA shape by how many CTAs of this footprint fit in an SM's 228 KB (+1 KB reserved each): three 320-thread CTAs
    // for the usual 30-waypoint trajectory, two of 512 threads for 50-60 waypoints, one of 1024 beyond, so that ~30
    // warps stay resident per SM either way.  OMGB_STEP_CONFIG = 0 / 1 / 2 / 3 forces 320x3 / 512x2 / 1024x1 / 256x4.
    static const int shape_threads[6] = {320, 512, 1024, 256, 192, 256}, shape_ctas[6] = {3, 2, 1, 4, 4, 3};
    auto layout_for = [&](int cfg) {
        return make_layout(n, c, lpi, s->num_objects, s->p, shape_threads[cfg] / 32, topk, fing);
    };
    auto fits = [&](int cfg) {
        const SmemLayout L = layout_for(cfg);
        return L.total <= (size_t)s->smem_optin && (size_t)shape_ctas[cfg] * (L.total + 1024u) <= 228u * 1024u;
    };
    int cfg = step_config();
    if (cfg < 0 || cfg > 5 || !fits(cfg)) {
        // (256x4 fits a 30-waypoint trajectory too; measured with a cold L2 it is slower than 320x3 -- 0.121 vs 0.117 ms
        // per step of 1024 trajectories: the per-trajectory critical path grows -- and faster with a warm one)
        cfg = (lpi == 16 && fits(0)) ? 0 : fits(1) ? 1 : 2;
        // small batches leave SMs under-filled: spend the idle warps on wider CTAs (lower latency per trajectory)
        if (a0.batch <= s->num_sms) cfg = 2;
        else if (a0.batch <= 2 * s->num_sms && (cfg == 0 || cfg == 3)) cfg = 1;
    }
    if (lpi == 32 && (cfg == 0 || cfg >= 3)) cfg = 1;
    const SmemLayout L = layout_for(cfg);
    if (L.total > (size_t)s->smem_optin)
        return fail(OMGB_ERR_UNSUPPORTED, "trajectory too long for one CTA's shared memory");
    StepArgs a = a_in;
    a.lay = L;
    a.rp = s->rp;
    if (topk) {   // the potential scratch covers pot_first + batch trajectories
        const size_t need = (pot_first + (size_t)a.batch) * (size_t)n * NL * lpi;
        if (need > s->pot_floats) {
            if (pot_first != 0) return fail(OMGB_ERR_STATE, "potential scratch not sized before a chunked launch");
            OMGB_CUDA(cudaStreamSynchronize(st));   // (earlier launches may still use the old buffer)
            if (s->has_last) OMGB_CUDA(cudaEventSynchronize(s->last_done));
            cudaFree(s->d_pot);
            s->d_pot = nullptr; s->pot_floats = 0;
            OMGB_CUDA(cudaMalloc(&s->d_pot, sizeof(float) * need));
            s->pot_floats = need;
        }
        a.pot_scratch = s->d_pot + pot_first * (size_t)n * NL * lpi;
    }
    if (guard_stream) { int g_ = stream_enter(s, st); if (g_) return g_; }
    // longest-first order from the previous launch on the same batch (same xi buffer and size)
    static int env_lpt = -1;
    if (env_lpt < 0) { const char *e = getenv("OMGB_NO_LPT"); env_lpt = (e && atoi(e)) ? 0 : 1; }
    OrderSlot *os = nullptr;
    if (!plan && env_lpt && s->use_lpt && a.batch >= 148) {
        for (int k = 0; k < ORDER_SLOTS; ++k)
            if (s->order[k].key == (const void *)a.xi && s->order[k].batch == a.batch) os = &s->order[k];
        if (!os) {
            os = &s->order[s->order_next];
            s->order_next = (s->order_next + 1) % ORDER_SLOTS;
            os->valid = false;
            os->age = 0;
            os->key = (const void *)a.xi;
            os->batch = a.batch;
        }
        if (a.batch > os->cap) {
            cudaFree(os->d_order); cudaFree(os->d_cost);
            os->d_order = os->d_cost = nullptr; os->cap = 0;
            OMGB_CUDA(cudaMalloc(&os->d_order, sizeof(int) * a.batch));
            OMGB_CUDA(cudaMalloc(&os->d_cost, sizeof(int) * a.batch));
            os->cap = a.batch;
            os->valid = false;
        }
        a.order = os->valid ? os->d_order : nullptr;
        a.cta_cost = os->d_cost;
    }
    int rc_ = OMGB_OK;
    if (lpi == 16) {
        if (cfg == 3) rc_ = launch_cfg<16, 256, 4>(a, L.total, st, plan, s->num_sms);
        else if (cfg == 4) rc_ = launch_cfg<16, 192, 4>(a, L.total, st, plan, s->num_sms);
        else if (cfg == 5) rc_ = launch_cfg<16, 256, 3>(a, L.total, st, plan, s->num_sms);
        else if (cfg == 0) rc_ = launch_cfg<16, 320, 3>(a, L.total, st, plan, s->num_sms);
        else if (cfg == 1) rc_ = launch_cfg<16, 512, 2>(a, L.total, st, plan, s->num_sms);
        else rc_ = launch_cfg<16, 1024, 1>(a, L.total, st, plan, s->num_sms);
    } else {
        if (cfg == 2) rc_ = launch_cfg<32, 1024, 1>(a, L.total, st, plan, s->num_sms);
        else rc_ = launch_cfg<32, 512, 2>(a, L.total, st, plan, s->num_sms);
    }
    if (rc_) return rc_;
    if (os && (!os->valid || (++os->age % LPT_REFRESH) == 0)) {
        // the per-trajectory cost changes slowly from one iteration to the next: re-sort every LPT_REFRESH launches
        lpt_order_kernel<<<1, 1024, 0, st>>>(os->d_cost, os->d_order, a.batch);
        ++g_launches;
        OMGB_CUDA(cudaGetLastError());
        os->valid = true;
    }
    if (guard_stream) return stream_leave(s, st);
    return OMGB_OK;
}

static StepArgs make_args(const omgb_scene *s, const omgb_step_params_t *prm, int batch, double *xi,
                          const double *start, const double *end, const double *goal_rows, const uint8_t *active,
                          double *grad_out, double *info, float *dbg_pot, float *dbg_pts, double *row_obs = nullptr) {
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.objs = s->d_objs; a.grids = s->d_grids; a.robot = s->d_robot; a.Ainv = s->d_Ainv; a.proj = s->d_proj;
    a.xi = xi; a.start = start; a.end = end; a.goal_rows = goal_rows; a.active = active; a.done = nullptr;
    a.grad_out = grad_out; a.info = info; a.dbg_pot = dbg_pot; a.dbg_pts = dbg_pts; a.row_obs = row_obs;
    a.dil = s->dil;
    a.quad = s->quad;
    a.prof = s->d_prof;
    a.num_objects = s->num_objects; a.batch = batch; a.iteration = 0; a.stop_on_terminate = 0;
    a.metric_kind = s->metric_kind; a.metric_scale = s->metric_scale;
    {   // TMA bulk staging of xi / object records (device pointers; the zero-copy host path switches it off)
        static int env_bulk = -1;
        if (env_bulk < 0) { const char *e = getenv("OMGB_NO_BULK"); env_bulk = (e && atoi(e)) ? 0 : 1; }
        a.bulk_stage = env_bulk;
    }
    {   // phase 4b: pairs one per lane from this many winners on (diagnostic override; results do not depend on it)
        static int env_seg = -1;
        if (env_seg < 0) { const char *e = getenv("OMGB_WIN_SEG_MIN"); env_seg = e ? atoi(e) : 0; }
        a.win_seg_min = env_seg;
    }
    a.prm = *prm;
    if (!a.prm.goal_set_proj) a.prm.constraint_rows = 0;
    return a;
}

extern "C" int omgb_chomp_step(omgb_scene_t *s, const omgb_step_params_t *prm, int batch, double *xi,
                               const double *start, const double *end, const double *goal_rows,
                               const uint8_t *active, double *grad_out, double *info, float *dbg_pot,
                               float *dbg_pts, double *row_obs, void *stream) {
    int rc_ = check_step(s, prm, batch, "omgb_chomp_step");
    if (rc_) return rc_;
    if (batch == 0) return OMGB_OK;   // (an empty tensor's data pointer is null)
    if (!xi || !start || !end || !info || (prm->goal_set_proj && !goal_rows))
        return fail(OMGB_ERR_INVALID, "omgb_chomp_step: null buffer");
    OMGB_CUDA(cudaSetDevice(s->device));
    StepArgs a = make_args(s, prm, batch, xi, start, end, goal_rows, active, grad_out, info, dbg_pot, dbg_pts, row_obs);
    return launch_step(s, a, (cudaStream_t)stream);
}

extern "C" int omgb_chomp_plan(omgb_scene_t *s, const omgb_step_params_t *prm, int iters, const double *ow,
                               const double *sw, const double *ss, int stop_on_terminate, int batch, double *xi,
                               const double *start, const double *end, const double *goal_rows, uint8_t *done,
                               double *info, void *stream) {
    return omgb_chomp_plan_history(s, prm, iters, ow, sw, ss, stop_on_terminate, batch, xi, start, end, goal_rows,
                                   done, info, nullptr, nullptr, stream);
}

extern "C" int omgb_chomp_plan_history(omgb_scene_t *s, const omgb_step_params_t *prm, int iters, const double *ow,
                                       const double *sw, const double *ss, int stop_on_terminate, int batch,
                                       double *xi, const double *start, const double *end, const double *goal_rows,
                                       uint8_t *done, double *info, double *hist_xi, double *hist_info,
                                       void *stream) {
    int rc_ = check_step(s, prm, batch, "omgb_chomp_plan");
    if (rc_) return rc_;
    if (iters < 0 || (iters > 0 && (!ow || !sw || !ss))) return fail(OMGB_ERR_INVALID, "omgb_chomp_plan: bad schedule");
    if (batch == 0 || iters == 0) return OMGB_OK;
    if (!xi || !start || !end || !info || (prm->goal_set_proj && !goal_rows) || (stop_on_terminate && !done))
        return fail(OMGB_ERR_INVALID, "omgb_chomp_plan: null buffer");
    OMGB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (done) OMGB_CUDA(cudaMemsetAsync(done, 0, (size_t)batch, st));
    StepArgs a = make_args(s, prm, batch, xi, start, end, goal_rows, nullptr, nullptr, info, nullptr, nullptr);
    a.done = stop_on_terminate ? done : nullptr;
    a.stop_on_terminate = stop_on_terminate;
    a.hist_xi = hist_xi; a.hist_info = hist_info;
    a.prm.update = 1;
    if (batch == 0 || iters == 0) return OMGB_OK;
    static int env_loop = -1;
    if (env_loop < 0) { const char *e = getenv("OMGB_PLAN_LAUNCHES"); env_loop = (e && atoi(e)) ? 1 : 0; }
    if (env_loop || s->d_prof) {
        // one launch per iteration (diagnostics / A-B; identical results)
        for (int it = 0; it < iters; ++it) {
            a.iteration = it;
            a.prm.obstacle_weight = ow[it];
            a.prm.smoothness_weight = sw[it];
            a.prm.step_size = ss[it];
            rc_ = launch_step(s, a, st);
            if (rc_) return rc_;
        }
        return OMGB_OK;
    }
    // persistent plan kernel: one launch, a device-side queue of (iteration, trajectory) items
    if (iters > s->plan_sched_cap) {
        cudaFree(s->d_plan_sched);
        s->d_plan_sched = nullptr; s->plan_sched_cap = 0;
        OMGB_CUDA(cudaMalloc(&s->d_plan_sched, sizeof(double) * 3 * iters));
        s->plan_sched_cap = iters;
    }
    if (batch > s->plan_progress_cap) {
        cudaFree(s->d_plan_progress);
        s->d_plan_progress = nullptr; s->plan_progress_cap = 0;
        OMGB_CUDA(cudaMalloc(&s->d_plan_progress, sizeof(int) * batch));
        s->plan_progress_cap = batch;
    }
    if (!s->d_plan_counter) OMGB_CUDA(cudaMalloc(&s->d_plan_counter, sizeof(unsigned)));
    if ((long long)batch * iters > 0x7fffffffLL) return fail(OMGB_ERR_INVALID, "omgb_chomp_plan: batch x iters too large");
    std::vector<double> sched(3 * (size_t)iters);
    for (int it = 0; it < iters; ++it) { sched[3 * it] = ow[it]; sched[3 * it + 1] = sw[it]; sched[3 * it + 2] = ss[it]; }
    OMGB_CUDA(cudaMemcpyAsync(s->d_plan_sched, sched.data(), sizeof(double) * 3 * iters, cudaMemcpyHostToDevice, st));
    OMGB_CUDA(cudaStreamSynchronize(st));   // (sched goes out of scope; pageable source)
    OMGB_CUDA(cudaMemsetAsync(s->d_plan_progress, 0, sizeof(int) * batch, st));
    OMGB_CUDA(cudaMemsetAsync(s->d_plan_counter, 0, sizeof(unsigned), st));
    PlanArgs pa;
    pa.sched = s->d_plan_sched; pa.progress = s->d_plan_progress; pa.counter = s->d_plan_counter; pa.iters = iters;
    return launch_step(s, a, st, &pa);
}

// Device alias of a host pointer when it lies in mapped pinned memory (cudaHostAlloc / cudaHostRegister under UVA).
static void *mapped_alias(const void *h) {
    if (!h) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost) return nullptr;
    return at.devicePointer;
}

extern "C" int omgb_chomp_step_host(omgb_scene_t *s, const omgb_step_params_t *prm, int batch, double *h_xi,
                                    const double *h_start, const double *h_end, const double *h_goal_rows,
                                    double *h_info, void *stream) {
    int rc_ = check_step(s, prm, batch, "omgb_chomp_step_host");
    if (rc_) return rc_;
    if (!h_xi || !h_start || !h_end || !h_info || (prm->goal_set_proj && !h_goal_rows))
        return fail(OMGB_ERR_INVALID, "omgb_chomp_step_host: null buffer");
    if (batch == 0) return OMGB_OK;
    OMGB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = prm->n_waypoints, c = prm->goal_set_proj ? prm->constraint_rows : 0;

    // (A) zero-copy: every buffer is mapped pinned host memory -> the fused kernel itself reads xi/start/end/goal rows
    // over PCIe while staging them into shared memory and writes the new xi and the info row straight back; there
    // is no separate copy to wait for.
    if (s->host_mode == 0 || s->host_mode == 3) {
        double *m_xi = (double *)mapped_alias(h_xi), *m_info = (double *)mapped_alias(h_info);
        const double *m_start = (const double *)mapped_alias(h_start), *m_end = (const double *)mapped_alias(h_end);
        const double *m_goal = c > 0 ? (const double *)mapped_alias(h_goal_rows) : nullptr;
        const bool all = m_xi && m_info && m_start && m_end && (c == 0 || m_goal);
        if (all) {
            StepArgs a = make_args(s, prm, batch, m_xi, m_start, m_end, m_goal, nullptr, nullptr, m_info, nullptr, nullptr);
            a.bulk_stage = 0;   // xi lives in mapped host memory: staged by per-thread loads issued back to back
            rc_ = launch_step(s, a, st);
            if (rc_) return rc_;
            OMGB_CUDA(cudaStreamSynchronize(st));
            return OMGB_OK;
        }
        if (s->host_mode == 3) return fail(OMGB_ERR_INVALID, "omgb_chomp_step_host: buffers are not mapped pinned memory");
    }

    // (B) staged: H2D -> kernel -> D2H, pipelined over up to PIPE_CHUNKS chunks of trajectories on the scene's own
    // streams so that chunk k's D2H overlaps chunk k+1's kernel and chunk k+2's H2D (separate copy engines).
    const size_t n_xi = (size_t)batch * n * ND, n_se = (size_t)batch * ND, n_goal = (size_t)batch * c * ND,
                 n_info = (size_t)batch * OMGB_INFO_STRIDE;
    const size_t need = sizeof(double) * (n_xi + 2 * n_se + n_goal + n_info);
    if (need > s->stage_bytes) {
        cudaFree(s->d_stage);
        s->d_stage = nullptr; s->stage_bytes = 0;
        OMGB_CUDA(cudaMalloc(&s->d_stage, need));
        s->stage_bytes = need;
    }
    double *d_xi = s->d_stage, *d_start = d_xi + n_xi, *d_end = d_start + n_se, *d_goal = d_end + n_se,
           *d_info = d_goal + n_goal;
    int chunks = 1;
    if (s->host_mode != 1) {
        chunks = batch / 256;
        chunks = chunks < 1 ? 1 : (chunks > PIPE_CHUNKS ? PIPE_CHUNKS : chunks);
    }
    if (chunks > 1 && !s->pipe_begin) {
        for (int k = 0; k < PIPE_CHUNKS; ++k) {
            OMGB_CUDA(cudaStreamCreateWithFlags(&s->pipe_stream[k], cudaStreamNonBlocking));
            OMGB_CUDA(cudaEventCreateWithFlags(&s->pipe_done[k], cudaEventDisableTiming));
        }
        OMGB_CUDA(cudaEventCreateWithFlags(&s->pipe_begin, cudaEventDisableTiming));
    }
    if (chunks > 1) {
        // size the potential scratch for the whole batch before the chunks start (they run concurrently)
        const size_t need = (size_t)batch * n * NL * (s->p <= 16 ? 16 : 32);
        if (prm->top_k_collision > 0 && need > s->pot_floats) {
            OMGB_CUDA(cudaStreamSynchronize(st));
            if (s->has_last) OMGB_CUDA(cudaEventSynchronize(s->last_done));
            cudaFree(s->d_pot);
            s->d_pot = nullptr; s->pot_floats = 0;
            OMGB_CUDA(cudaMalloc(&s->d_pot, sizeof(float) * need));
            s->pot_floats = need;
        }
        rc_ = stream_enter(s, st);
        if (rc_) return rc_;
        OMGB_CUDA(cudaEventRecord(s->pipe_begin, st));
    }
    for (int k = 0; k < chunks; ++k) {
        const int b0 = (int)((long long)batch * k / chunks), b1 = (int)((long long)batch * (k + 1) / chunks);
        const size_t nb = (size_t)(b1 - b0);
        cudaStream_t cs = chunks > 1 ? s->pipe_stream[k] : st;
        if (chunks > 1) OMGB_CUDA(cudaStreamWaitEvent(cs, s->pipe_begin, 0));
        OMGB_CUDA(cudaMemcpyAsync(d_xi + (size_t)b0 * n * ND, h_xi + (size_t)b0 * n * ND, sizeof(double) * nb * n * ND,
                                  cudaMemcpyHostToDevice, cs));
        OMGB_CUDA(cudaMemcpyAsync(d_start + (size_t)b0 * ND, h_start + (size_t)b0 * ND, sizeof(double) * nb * ND,
                                  cudaMemcpyHostToDevice, cs));
        OMGB_CUDA(cudaMemcpyAsync(d_end + (size_t)b0 * ND, h_end + (size_t)b0 * ND, sizeof(double) * nb * ND,
                                  cudaMemcpyHostToDevice, cs));
        if (c > 0)
            OMGB_CUDA(cudaMemcpyAsync(d_goal + (size_t)b0 * c * ND, h_goal_rows + (size_t)b0 * c * ND,
                                      sizeof(double) * nb * c * ND, cudaMemcpyHostToDevice, cs));
        StepArgs a = make_args(s, prm, (int)nb, d_xi + (size_t)b0 * n * ND, d_start + (size_t)b0 * ND,
                               d_end + (size_t)b0 * ND, c > 0 ? d_goal + (size_t)b0 * c * ND : nullptr, nullptr, nullptr,
                               d_info + (size_t)b0 * OMGB_INFO_STRIDE, nullptr, nullptr);
        rc_ = chunks > 1 ? launch_step(s, a, cs, nullptr, (size_t)b0, false) : launch_step(s, a, cs);
        if (rc_) return rc_;
        OMGB_CUDA(cudaMemcpyAsync(h_xi + (size_t)b0 * n * ND, d_xi + (size_t)b0 * n * ND, sizeof(double) * nb * n * ND,
                                  cudaMemcpyDeviceToHost, cs));
        OMGB_CUDA(cudaMemcpyAsync(h_info + (size_t)b0 * OMGB_INFO_STRIDE, d_info + (size_t)b0 * OMGB_INFO_STRIDE,
                                  sizeof(double) * nb * OMGB_INFO_STRIDE, cudaMemcpyDeviceToHost, cs));
        if (chunks > 1) {
            OMGB_CUDA(cudaEventRecord(s->pipe_done[k], cs));
            OMGB_CUDA(cudaStreamWaitEvent(st, s->pipe_done[k], 0));
        }
    }
    OMGB_CUDA(cudaStreamSynchronize(st));
    return OMGB_OK;
}

extern "C" int omgb_batch_obstacle_cost(omgb_scene_t *s, const double *joints, int num_configs, int arc_length,
                                        const double *start, double time_interval, int uncheck_finger_collision,
                                        float *potentials, float *grads, float *collides, void *stream) {
    if (!s) return fail(OMGB_ERR_INVALID, "omgb_batch_obstacle_cost: null scene");
    if (!s->robot_set || !s->sdf_set || !s->objs_set)
        return fail(OMGB_ERR_STATE, "omgb_batch_obstacle_cost: scene needs robot, sdf and objects");
    if (num_configs < 0) return fail(OMGB_ERR_INVALID, "omgb_batch_obstacle_cost: negative num_configs");
    if (num_configs == 0) return OMGB_OK;
    if (!joints || !potentials || !collides) return fail(OMGB_ERR_INVALID, "omgb_batch_obstacle_cost: null buffer");
    if (arc_length > 0 && (!start || num_configs % arc_length != 0 || !(time_interval > 0)))
        return fail(OMGB_ERR_INVALID, "omgb_batch_obstacle_cost: arc_length needs start, dt and M % arc_length == 0");
    OMGB_CUDA(cudaSetDevice(s->device));
    const size_t smem = align_up(sizeof(ObjRec) * s->num_objects, 16) + sizeof(double) * (BOC_CFG + 2) * NL * 12;
    OMGB_CUDA(cudaFuncSetAttribute(batch_obstacle_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (num_configs + BOC_CFG - 1) / BOC_CFG;
    batch_obstacle_cost_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(
        s->d_objs, s->num_objects, s->d_grids, s->d_robot, joints, num_configs, arc_length, start,
        arc_length > 0 ? (float)(1.0 / time_interval) : 0.0f, uncheck_finger_collision == -1 ? 1 : 0, potentials,
        grads, collides);
    OMGB_CUDA(cudaGetLastError());
    return OMGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// goal scoring (device half of Learner.cost_vector)
// ----------------------------------------------------------------------------------------------------
extern "C" int omgb_goal_costs(omgb_scene_t *s, int batch, const double *from, long long from_stride,
                               const double *goals, int num_goals, int goals_shared, int arc_length,
                               double time_interval, int uncheck_finger_collision, float *costs, void *stream) {
    if (!s) return fail(OMGB_ERR_INVALID, "omgb_goal_costs: null scene");
    if (!s->robot_set || !s->sdf_set || !s->objs_set)
        return fail(OMGB_ERR_STATE, "omgb_goal_costs: scene needs robot, sdf and objects");
    if (batch < 0 || num_goals < 0) return fail(OMGB_ERR_INVALID, "omgb_goal_costs: negative size");
    if (batch == 0 || num_goals == 0) return OMGB_OK;
    if (!from || !goals || !costs) return fail(OMGB_ERR_INVALID, "omgb_goal_costs: null buffer");
    if (arc_length < 1 || !(time_interval > 0) || from_stride < ND)
        return fail(OMGB_ERR_INVALID, "omgb_goal_costs: arc_length >= 1, time_interval > 0 and from_stride >= 9 required");
    if ((long long)batch * num_goals > 0x7fffffffLL) return fail(OMGB_ERR_INVALID, "omgb_goal_costs: batch x goals too large");
    OMGB_CUDA(cudaSetDevice(s->device));
    GoalArgs a;
    memset(&a, 0, sizeof(a));
    a.objs = s->d_objs; a.grids = s->d_grids; a.robot = s->d_robot;
    a.quad = s->quad;
    a.from = from; a.from_stride = from_stride;
    a.goals = goals; a.goal_stride_b = goals_shared ? 0 : (long long)num_goals * ND;
    a.costs = costs;
    a.dil = s->dil;
    a.rp = s->rp;
    a.num_objects = s->num_objects; a.num_goals = num_goals; a.arc = arc_length;
    a.finger_soft = uncheck_finger_collision == -1 ? 1 : 0;
    a.inv_dt = (float)(1.0 / time_interval);
    // goals per CTA: as many lines as fit ~30 configurations; block size by the number of link instances of a CTA
    // (goals x arc x 10): the cull is one instance per thread
    int gpc = arc_length >= 16 ? 1 : 30 / arc_length;
    gpc = gpc > num_goals ? num_goals : (gpc > GOAL_MAX_GPC ? GOAL_MAX_GPC : gpc);
    {
        static int env_gpc = -1;
        if (env_gpc < 0) { const char *e = getenv("OMGB_GOAL_GPC"); env_gpc = e ? atoi(e) : 0; }
        if (env_gpc > 0) gpc = env_gpc > num_goals ? num_goals : (env_gpc > GOAL_MAX_GPC ? GOAL_MAX_GPC : env_gpc);
    }
    a.gpc = gpc;
    a.ctas_per_traj = (num_goals + gpc - 1) / gpc;
    const int n_li = gpc * arc_length * NL;
    const int lpi = s->p <= 16 ? 16 : 32;
    const int shape = lpi == 32 ? 2 : (n_li <= 128 ? 0 : n_li <= 192 ? 1 : n_li <= 256 ? 2 : 3);
    static const int shape_threads[4] = {128, 192, 256, 320};
    const int hi = s->num_objects > 32 ? 1 : 0;
    goal_layout(a, shape_threads[shape] / 32);
    if (a.smem_total > (unsigned)s->smem_optin)
        return fail(OMGB_ERR_UNSUPPORTED, "omgb_goal_costs: arc_length too long for one CTA's shared memory");
    if ((long long)batch * a.ctas_per_traj > 0x7fffffffLL) return fail(OMGB_ERR_INVALID, "omgb_goal_costs: batch x goals too large");
    const int grid = batch * a.ctas_per_traj;
    cudaStream_t gst = (cudaStream_t)stream;
    static unsigned cached[2][4][2][64] = {{{{0}}}};
    cudaError_t e_ = cudaSuccess;
    {
        std::lock_guard<std::mutex> lock(g_attr_mutex);
        const bool set_attr = s->device >= 64 || cached[lpi == 32][shape][hi][s->device] < a.smem_total;
#define OMGB_GOAL_CASE(T, L, H)                                                                                        \
    do {                                                                                                               \
        if (set_attr) e_ = cudaFuncSetAttribute(goal_cost_kernel<T, L, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                (int)a.smem_total);                                                    \
        if (e_ == cudaSuccess) goal_cost_kernel<T, L, H><<<grid, T, a.smem_total, gst>>>(a);                           \
    } while (0)
        if (lpi == 32) { if (hi) OMGB_GOAL_CASE(256, 32, true); else OMGB_GOAL_CASE(256, 32, false); }
        else if (shape == 0) { if (hi) OMGB_GOAL_CASE(128, 16, true); else OMGB_GOAL_CASE(128, 16, false); }
        else if (shape == 1) { if (hi) OMGB_GOAL_CASE(192, 16, true); else OMGB_GOAL_CASE(192, 16, false); }
        else if (shape == 2) { if (hi) OMGB_GOAL_CASE(256, 16, true); else OMGB_GOAL_CASE(256, 16, false); }
        else { if (hi) OMGB_GOAL_CASE(320, 16, true); else OMGB_GOAL_CASE(320, 16, false); }
#undef OMGB_GOAL_CASE
        if (e_ != cudaSuccess) return fail(OMGB_ERR_CUDA, std::string("goal_cost_kernel attributes: ") + cudaGetErrorString(e_));
        if (set_attr && s->device < 64) cached[lpi == 32][shape][hi][s->device] = a.smem_total;
    }
    ++g_launches;
    OMGB_CUDA(cudaGetLastError());
    return OMGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// trajectory initialisation (omg/util.py:238-258) and the SDF asset path (omg/core.py:366-457); these run on
// the CURRENT device of the calling thread and need no scene
// ----------------------------------------------------------------------------------------------------
static int grid_for(long long work_items, int threads) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)sms * 16;   // grid-stride kernels: a multiple of the SM count
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

extern "C" int omgb_traj_interpolate(const double *waypoints, int batch, int num_knots, int n_waypoints, int mode,
                                     double *xi, void *stream) {
    if (batch < 0 || n_waypoints < 0) return fail(OMGB_ERR_INVALID, "omgb_traj_interpolate: negative size");
    if (num_knots < 2 || num_knots > TRAJ_MAX_KNOTS)
        return fail(OMGB_ERR_INVALID, "omgb_traj_interpolate: 2 <= num_knots <= 32 required");
    if (mode != 0 && mode != 1) return fail(OMGB_ERR_INVALID, "omgb_traj_interpolate: mode is 0 (linear) or 1 (cubic)");
    if (batch == 0 || n_waypoints == 0) return OMGB_OK;
    if (!waypoints || !xi) return fail(OMGB_ERR_INVALID, "omgb_traj_interpolate: null buffer");
    const long long total = (long long)batch * n_waypoints * ND;
    traj_interpolate_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(waypoints, batch, num_knots,
                                                                                    n_waypoints, mode, xi);
    ++g_launches;
    OMGB_CUDA(cudaGetLastError());
    return OMGB_OK;
}

struct SdfSourceTable {
    SdfSource src[OMGB_MAX_OBJECTS];
};

__global__ void __launch_bounds__(256) sdf_pack_entry(const __grid_constant__ SdfSourceTable tab, int num_objects, int X,
                                                      int Y, int Z, float *__restrict__ dst);

extern "C" int omgb_sdf_pack(const omgb_sdf_source_t *sources, int num_objects, int dim_x, int dim_y, int dim_z,
                             float *d_out, void *stream) {
    if (num_objects < 0 || num_objects > OMGB_MAX_OBJECTS)
        return fail(OMGB_ERR_INVALID, "omgb_sdf_pack: 0 <= num_objects <= 64 required");
    if (num_objects == 0) return OMGB_OK;
    if (!sources || !d_out || dim_x <= 0 || dim_y <= 0 || dim_z <= 0)
        return fail(OMGB_ERR_INVALID, "omgb_sdf_pack: null buffer or empty shape");
    SdfSourceTable tab;
    memset(&tab, 0, sizeof(tab));
    for (int o = 0; o < num_objects; ++o) {
        const omgb_sdf_source_t &q = sources[o];
        if (!q.data || q.shape[0] <= 0 || q.shape[1] <= 0 || q.shape[2] <= 0 || q.shape[0] > dim_x ||
            q.shape[1] > dim_y || q.shape[2] > dim_z)
            return fail(OMGB_ERR_INVALID, "omgb_sdf_pack: object grid missing or larger than the padded shape");
        if ((q.layout != 0 && q.layout != 1) || (q.dtype != 0 && q.dtype != 1))
            return fail(OMGB_ERR_INVALID, "omgb_sdf_pack: layout and dtype are 0 or 1");
        tab.src[o].data = q.data;
        tab.src[o].sx = q.shape[0]; tab.src[o].sy = q.shape[1]; tab.src[o].sz = q.shape[2];
        tab.src[o].layout = q.layout; tab.src[o].dtype = q.dtype; tab.src[o].scale = q.scale;
    }
    if ((long long)num_objects * dim_x > 65535LL * 1024)
        return fail(OMGB_ERR_INVALID, "omgb_sdf_pack: too many (object, x) planes");
    const int plane = dim_y * ((dim_z + 3) / 4);
    const long long planes = (long long)num_objects * dim_x;
    if (planes > 0x7fffffffLL) return fail(OMGB_ERR_INVALID, "omgb_sdf_pack: shape too large");
    int bx = (plane + 1023) / 1024;   // ~4 groups (64 B) per thread
    if (bx < 1) bx = 1;
    // grid.y is limited to 65535: fold the rest into z
    dim3 grid((unsigned)bx, (unsigned)(planes > 65535 ? 65535 : planes), (unsigned)((planes + 65534) / 65535));
    sdf_pack_entry<<<grid, 256, 0, (cudaStream_t)stream>>>(tab, num_objects, dim_x, dim_y, dim_z, d_out);
    ++g_launches;
    OMGB_CUDA(cudaGetLastError());
    return OMGB_OK;
}

__global__ void __launch_bounds__(256) sdf_pack_entry(const __grid_constant__ SdfSourceTable tab, int num_objects, int X,
                                                      int Y, int Z, float *__restrict__ dst) {
    sdf_pack_body(tab.src, num_objects, X, Y, Z, dst);
}

extern "C" int omgb_point_sdf(const double *d_points, int num_points, const double *d_gx, const double *d_gy,
                              const double *d_gz, int dim_x, int dim_y, int dim_z, float *d_out32, double *d_out64,
                              void *stream) {
    if (num_points < 1) return fail(OMGB_ERR_INVALID, "omgb_point_sdf: at least one point required");
    if (dim_x < 0 || dim_y < 0 || dim_z < 0) return fail(OMGB_ERR_INVALID, "omgb_point_sdf: negative shape");
    const long long total = (long long)dim_x * dim_y * dim_z;
    if (total == 0) return OMGB_OK;
    if (!d_points || !d_gx || !d_gy || !d_gz || (!d_out32 && !d_out64))
        return fail(OMGB_ERR_INVALID, "omgb_point_sdf: null buffer");
    const long long items = (long long)dim_x * dim_y * ((dim_z + POINT_ZV - 1) / POINT_ZV);   // (x, y, z-chunk)
    const long long blocks = (items + 31) / 32;   // 32 work items per block, POINT_SLICES warps each
    if (blocks > 0x7fffffffLL) return fail(OMGB_ERR_INVALID, "omgb_point_sdf: grid too large");
    point_sdf_kernel<<<(int)blocks, POINT_THREADS, 0, (cudaStream_t)stream>>>(d_points, num_points, d_gx, d_gy, d_gz,
                                                                             dim_x, dim_y, dim_z, d_out32, d_out64);
    ++g_launches;
    OMGB_CUDA(cudaGetLastError());
    return OMGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// goal-set plans with goal switching, device resident: one plan iteration + the learner's update
// ----------------------------------------------------------------------------------------------------
extern "C" int omgb_chomp_plan_step(omgb_scene_t *s, const omgb_step_params_t *prm, int iteration, int stop_on_terminate,
                                    int batch, double *xi, const double *start, const double *end,
                                    const double *goal_rows, uint8_t *done, double *info, double *hist_xi,
                                    double *hist_info, void *stream) {
    int rc_ = check_step(s, prm, batch, "omgb_chomp_plan_step");
    if (rc_) return rc_;
    if (iteration < 0) return fail(OMGB_ERR_INVALID, "omgb_chomp_plan_step: negative iteration");
    if (batch == 0) return OMGB_OK;
    if (!xi || !start || !end || !info || (prm->goal_set_proj && !goal_rows) || (stop_on_terminate && !done))
        return fail(OMGB_ERR_INVALID, "omgb_chomp_plan_step: null buffer");
    OMGB_CUDA(cudaSetDevice(s->device));
    StepArgs a = make_args(s, prm, batch, xi, start, end, goal_rows, nullptr, nullptr, info, nullptr, nullptr);
    a.done = done;
    a.stop_on_terminate = stop_on_terminate;
    a.iteration = iteration;
    a.hist_xi = hist_xi; a.hist_info = hist_info;
    a.prm.update = 1;
    return launch_step(s, a, (cudaStream_t)stream);
}

extern "C" int omgb_learner_update(const omgb_learner_params_t *prm, int batch, const double *xi,
                                   const float *collision, const double *goal_set, int goals_shared,
                                   const double *reach, double *p, double *sum_costs, double *experts_p,
                                   double *experts_costs, double *q, const uint8_t *done, int32_t *goal_idx,
                                   double *end, double *goal_rows, double *cost_vector, int32_t *selected,
                                   void *stream) {
    if (!prm) return fail(OMGB_ERR_INVALID, "omgb_learner_update: null params");
    if (batch < 0) return fail(OMGB_ERR_INVALID, "omgb_learner_update: negative batch");
    if (prm->alg < OMGB_LEARNER_FTL || prm->alg > OMGB_LEARNER_INIT)
        return fail(OMGB_ERR_INVALID, "omgb_learner_update: unknown algorithm");
    if (prm->num_goals < 1 || prm->num_goals > LRN_MAX_GOALS)
        return fail(OMGB_ERR_INVALID, "omgb_learner_update: 1 <= num_goals <= 256 required");
    if (prm->n_waypoints < 1 || prm->first_waypoint < 0 || prm->first_waypoint >= prm->n_waypoints ||
        prm->constraint_rows < 1)
        return fail(OMGB_ERR_INVALID, "omgb_learner_update: bad waypoint / constraint_rows");
    if (batch == 0) return OMGB_OK;
    if (!xi || !goal_set || !goal_idx || !end || !p || (prm->alg != OMGB_LEARNER_PROJ && !collision))
        return fail(OMGB_ERR_INVALID, "omgb_learner_update: null buffer");
    if ((prm->alg == OMGB_LEARNER_FTL || prm->alg == OMGB_LEARNER_EXP) && !sum_costs)
        return fail(OMGB_ERR_INVALID, "omgb_learner_update: sum_costs required");
    if (prm->alg == OMGB_LEARNER_MD && (!experts_p || !experts_costs || !q))
        return fail(OMGB_ERR_INVALID, "omgb_learner_update: expert state required");
    LearnerArgs a;
    memset(&a, 0, sizeof(a));
    a.prm = *prm;
    a.xi = xi; a.collision = collision;
    a.goal_set = goal_set; a.goal_stride_b = goals_shared ? 0 : (long long)prm->num_goals * ND;
    a.reach = reach; a.reach_stride_b = goals_shared ? 0 : (long long)prm->num_goals * prm->constraint_rows * ND;
    a.p = p; a.sum_costs = sum_costs; a.experts_p = experts_p; a.experts_costs = experts_costs; a.q = q;
    a.done = done; a.goal_idx = goal_idx; a.end = end; a.goal_rows = goal_rows; a.cost_vector = cost_vector;
    a.selected_hist = selected; a.batch = batch;
    const int G = prm->num_goals;
    cudaStream_t st = (cudaStream_t)stream;
    // Few trajectories: the launch lasts as long as one warp's bisection chain -> the two-steps-per-round form
    // (identical results, learner_bisect.h); many: the SMs are full, the extra evaluations would only cost throughput.
    static int env_two = -1;
    if (env_two < 0) { const char *e = getenv("OMGB_LEARNER_TWO_STEP"); env_two = e ? (atoi(e) ? 1 : 0) : 2; }
    const bool two = prm->alg == OMGB_LEARNER_MD && (env_two == 2 ? batch <= 2 * 148 : env_two == 1);
    if (G <= 32) {
        if (two) learner_update_kernel<1, true><<<batch, LRN_THREADS, 0, st>>>(a);
        else learner_update_kernel<1><<<batch, LRN_THREADS, 0, st>>>(a);
    } else if (G <= 64) {
        if (two) learner_update_kernel<2, true><<<batch, LRN_THREADS, 0, st>>>(a);
        else learner_update_kernel<2><<<batch, LRN_THREADS, 0, st>>>(a);
    }
    else if (G <= 128) learner_update_kernel<4><<<batch, LRN_THREADS, 0, st>>>(a);
    else learner_update_kernel<8><<<batch, LRN_THREADS, 0, st>>>(a);
    ++g_launches;
    OMGB_CUDA(cudaGetLastError());
    return OMGB_OK;
}

// ----------------------------------------------------------------------------------------------------
// A whole goal-set plan with goal switching enqueued by ONE call: per iteration goal scoring -> learner update ->
// plan step, exactly the three entry points above with the arguments the host mirror passed them one call at a time
// (omg/planner.py:612-635).  The loop runs here instead of in the caller's interpreter: at one trajectory (the
// reference's own shape) the three launches of an iteration take less GPU time than the Python around them.
// ----------------------------------------------------------------------------------------------------
extern "C" int omgb_chomp_plan_goalset(omgb_scene_t *s, const omgb_step_params_t *prm,
                                       const omgb_learner_params_t *learner, int iters, int learner_iters,
                                       const double *schedule, const int32_t *first_waypoint, int batch,
                                       const omgb_goalset_plan_buffers_t *buf, double timeout_s, int *iters_enqueued,
                                       void *stream) {
    if (iters_enqueued) *iters_enqueued = 0;
    int rc_ = check_step(s, prm, batch, "omgb_chomp_plan_goalset");
    if (rc_) return rc_;
    if (!learner || !buf) return fail(OMGB_ERR_INVALID, "omgb_chomp_plan_goalset: null params");
    if (iters < 0 || learner_iters < 0 || learner_iters > iters)
        return fail(OMGB_ERR_INVALID, "omgb_chomp_plan_goalset: 0 <= learner_iters <= iters required");
    if (!prm->goal_set_proj) return fail(OMGB_ERR_INVALID, "omgb_chomp_plan_goalset: goal_set_proj required");
    if (batch == 0 || iters == 0) return OMGB_OK;
    if (!schedule || (learner_iters > 0 && !first_waypoint))
        return fail(OMGB_ERR_INVALID, "omgb_chomp_plan_goalset: null schedule");
    if (!buf->xi || !buf->start || !buf->end || !buf->goal_rows || !buf->done || !buf->info || !buf->goal_set ||
        !buf->reach_goals || !buf->goal_idx)
        return fail(OMGB_ERR_INVALID, "omgb_chomp_plan_goalset: null buffer");
    const bool score = learner->alg != OMGB_LEARNER_PROJ;
    if (learner_iters > 0 && score && !buf->collision)
        return fail(OMGB_ERR_INVALID, "omgb_chomp_plan_goalset: collision scratch required");
    const int n = prm->n_waypoints;
    omgb_step_params_t sp = *prm;
    omgb_learner_params_t lp = *learner;
    // cfg.timeout (omg/planner.py:629) against the DEVICE's progress: every 8 iterations an event is recorded; before
    // the clock is read the event two checks back has completed, so at most 16 iterations are in flight past it
    const bool timed = timeout_s >= 0.0;
    const auto t_begin = std::chrono::steady_clock::now();
    std::vector<cudaEvent_t> checks;
    auto drop_checks = [&]() { for (cudaEvent_t e : checks) cudaEventDestroy(e); checks.clear(); };
    cudaStream_t st = (cudaStream_t)stream;
    for (int t = 0; t < iters; ++t) {
        if (timed && t > 0 && t % 8 == 0) {
            if (checks.size() >= 2) cudaEventSynchronize(checks[checks.size() - 2]);
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
            if (el > timeout_s) break;
            cudaEvent_t e;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess) {
                cudaEventRecord(e, st);
                checks.push_back(e);
            }
        }
        if (t < learner_iters) {
            const int first = first_waypoint[t];
            if (first < 0 || first >= n) { drop_checks(); return fail(OMGB_ERR_INVALID, "omgb_chomp_plan_goalset: first waypoint out of range"); }
            lp.first_waypoint = first;
            if (score) {
                rc_ = omgb_goal_costs(s, batch, buf->xi + (size_t)ND * first, (long long)n * ND, buf->reach_goals,
                                      lp.num_goals, buf->goals_shared, n - first, prm->time_interval, 0, buf->collision,
                                      stream);
                if (rc_) { drop_checks(); return rc_; }
            }
            rc_ = omgb_learner_update(&lp, batch, buf->xi, score ? buf->collision : nullptr, buf->goal_set,
                                      buf->goals_shared, buf->reach, buf->p, buf->sum_costs, buf->experts_p,
                                      buf->experts_costs, buf->q, buf->done, buf->goal_idx, buf->end, buf->goal_rows,
                                      nullptr, buf->selected ? buf->selected + (size_t)t * batch : nullptr, stream);
            if (rc_) { drop_checks(); return rc_; }
        }
        sp.obstacle_weight = schedule[3 * t];
        sp.smoothness_weight = schedule[3 * t + 1];
        sp.step_size = schedule[3 * t + 2];
        rc_ = omgb_chomp_plan_step(s, &sp, t, 1, batch, buf->xi, buf->start, buf->end, buf->goal_rows, buf->done,
                                   buf->info, buf->hist_xi, buf->hist_info, stream);
        if (rc_) { drop_checks(); return rc_; }
        if (iters_enqueued) *iters_enqueued = t + 1;
    }
    drop_checks();
    return OMGB_OK;
}
