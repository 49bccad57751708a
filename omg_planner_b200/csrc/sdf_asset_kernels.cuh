// SDF asset path on the device.
//
//  sdf_pack_kernel        Env.combine_sdfs (omg/core.py:366-411) fused with SignedDensityField.from_pth's
//                         permute(1,0,2) (omg/sdf_tools.py:187-193) and SignedDensityField.resize's `data *= ratio`
//                         (omg/sdf_tools.py:37-39): every object's raw grid is read once, in whatever layout it
//                         has on disk, and written once into its slot of the padded [O,X,Y,Z] fp32 tensor (pad value
//                         1.0).  One launch for the whole scene; HBM-bound: O*X*Y*Z*4 bytes written, the raw grids read.
//  point_sdf_kernel       PointEnv.compute_sdf_from_points (omg/core.py:426-457): distance from every voxel of the
//                         workspace grid to the nearest point of a point cloud (the reference: scipy cKDTree.query).
//                         Brute force with the cloud staged through shared memory; fp64 with the exact operation
//                         order of cKDTree's squared-distance loop ((dx^2 + dy^2) + dz^2, no contraction), so the
//                         result is bit-identical.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace omgb {

struct SdfSource {          // one object's raw grid (DEVICE pointer); the table travels as a kernel parameter
    const void *data;
    int sx, sy, sz;         // logical shape [X,Y,Z] of the object's SDF (after from_pth's permute)
    int layout;             // 0: stored [X,Y,Z]; 1: stored [Y,X,Z] (the .pth files: sdf_torch[0,0], convert_sdf.py:43)
    int dtype;              // 0: fp32, 1: fp64 (PointEnv hands fp64 distances to SignedDensityField)
    float scale;            // SignedDensityField.resize ratio applied to data_torch (fp32 multiply); 1 = none
};

__device__ __forceinline__ float sdf_source_at(const SdfSource &s, int x, int y, int z) {
    const size_t idx = s.layout ? ((size_t)y * s.sx + x) * s.sz + z : ((size_t)x * s.sy + y) * s.sz + z;
    float v = s.dtype ? (float)__ldg(reinterpret_cast<const double *>(s.data) + idx)
                      : __ldg(reinterpret_cast<const float *>(s.data) + idx);
    if (s.scale != 1.0f) v = __fmul_rn(v, s.scale);
    return v;
}

// blockIdx.y = (object, x) plane of the padded tensor; threads grid-stride over the plane's (y, z/4) groups: one
// 128-bit store each when Z % 4 == 0, and one 128-bit load when the source row allows it (both layouts keep z
// contiguous), so the kernel moves every byte exactly once with full-width transactions.  No 64-bit divisions.
__device__ __forceinline__ void sdf_pack_body(const SdfSource *__restrict__ src, int num_objects, int X, int Y, int Z,
                                              float *__restrict__ dst) {
    const int zq = (Z + 3) >> 2;
    const int plane = Y * zq;
    const long long pl = (long long)blockIdx.z * 65535 + blockIdx.y;   // (object, x) plane
    const int o = (int)(pl / X), x = (int)(pl - (long long)o * X);
    if (o >= num_objects) return;
    const SdfSource s = src[o];
    const bool vec_st = (Z & 3) == 0;
    const bool vec_ld = s.dtype == 0 && (s.sz & 3) == 0 && ((reinterpret_cast<size_t>(s.data) & 15) == 0);
    float *plane_out = dst + ((size_t)o * X + x) * (size_t)Y * Z;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < plane; g += gridDim.x * blockDim.x) {
        const int y = g / zq, z0 = (g - y * zq) * 4;
        float v[4];
        const bool row_in = x < s.sx && y < s.sy;
        if (row_in && vec_ld && z0 + 3 < s.sz) {
            const size_t idx = s.layout ? ((size_t)y * s.sx + x) * s.sz + z0 : ((size_t)x * s.sy + y) * s.sz + z0;
            const float4 t = __ldcs(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(s.data) + idx));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            if (s.scale != 1.0f) {
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = __fmul_rn(v[q], s.scale);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int z = z0 + q;
                v[q] = (row_in && z < s.sz) ? sdf_source_at(s, x, y, z) : 1.0f;
            }
        }
        float *out = plane_out + (size_t)y * Z + z0;
        if (vec_st) {
            __stcs(reinterpret_cast<float4 *>(out), make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (z0 + q < Z) out[q] = v[q];
        }
    }
}

constexpr int POINT_TILE = 1024;
constexpr int POINT_ZV = 8;          // voxels of one z-row per thread
constexpr int POINT_SLICES = 4;      // warps of a block share 32 voxel chunks and split the cloud
constexpr int POINT_THREADS = 32 * POINT_SLICES;

// Work item = (x, y, chunk of POINT_ZV voxels along z), one per LANE; the block's POINT_SLICES warps all work on the
// same 32 items and each scans every POINT_SLICES-th point of the cloud (a workspace grid has only ~10^5 voxels: the
// split over points is what fills 148 SMs).  The cloud streams through shared memory in tiles; all lanes of a warp
// read the same point (broadcast).  The voxels of a lane share x and y, so dx*dx + dy*dy -- the first two terms of
// cKDTree's sum -- is computed once per point and only (.. + dz*dz) per voxel: the same IEEE operations per
// (voxel, point) as the reference, 4.6 instead of 9 fp64 instructions.  min is exact, so the split changes nothing.
__global__ void __launch_bounds__(POINT_THREADS) point_sdf_kernel(const double *__restrict__ points, int num_points,
                                                                  const double *__restrict__ gx,
                                                                  const double *__restrict__ gy,
                                                                  const double *__restrict__ gz, int X, int Y, int Z,
                                                                  float *__restrict__ out32,
                                                                  double *__restrict__ out64) {
    __shared__ double s_p[POINT_TILE * 3];
    __shared__ double s_best[POINT_SLICES][POINT_ZV][32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int chunks = (Z + POINT_ZV - 1) / POINT_ZV;
    const long long items = (long long)X * Y * chunks;
    const long long it = (long long)blockIdx.x * 32 + lane;
    const bool live = it < items;
    double px = 0.0, py = 0.0, pz[POINT_ZV], best[POINT_ZV];
    int z0 = 0;
    long long row = 0;
    if (live) {
        const int ch = (int)(it % chunks);
        row = it / chunks;
        z0 = ch * POINT_ZV;
        px = __ldg(gx + (int)(row / Y));
        py = __ldg(gy + (int)(row % Y));
    }
#pragma unroll
    for (int v = 0; v < POINT_ZV; ++v) {
        pz[v] = (live && z0 + v < Z) ? __ldg(gz + z0 + v) : 0.0;
        best[v] = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    }
    for (int t0 = 0; t0 < num_points; t0 += POINT_TILE) {
        const int cnt = min(POINT_TILE, num_points - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt * 3; k += blockDim.x) s_p[k] = __ldg(points + (size_t)t0 * 3 + k);
        __syncthreads();
        if (live) {
#pragma unroll 2
            for (int k = slice; k < cnt; k += POINT_SLICES) {
                const double qx = s_p[3 * k], qy = s_p[3 * k + 1], qz = s_p[3 * k + 2];
                const double dx = __dsub_rn(px, qx), dy = __dsub_rn(py, qy);
                const double dxy = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
#pragma unroll
                for (int v = 0; v < POINT_ZV; ++v) {
                    const double dz = __dsub_rn(pz[v], qz);
                    best[v] = fmin(best[v], __dadd_rn(dxy, __dmul_rn(dz, dz)));
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < POINT_ZV; ++v) s_best[slice][v][lane] = best[v];
    __syncthreads();
    // combine the slices: thread (slice, lane) finishes voxels v = slice, slice + POINT_SLICES, ... of lane's chunk
    if (live) {
        for (int v = slice; v < POINT_ZV; v += POINT_SLICES) {
            if (z0 + v >= Z) continue;
            double m = s_best[0][v][lane];
#pragma unroll
            for (int q = 1; q < POINT_SLICES; ++q) m = fmin(m, s_best[q][v][lane]);
            const double d = __dsqrt_rn(m);
            const size_t o = (size_t)row * Z + z0 + v;
            if (out64) out64[o] = d;
            if (out32) out32[o] = (float)d;
        }
    }
}

}  // namespace omgb
