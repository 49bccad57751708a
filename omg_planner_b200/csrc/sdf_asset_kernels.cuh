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

// grid-stride over groups of 4 consecutive z (one 128-bit store each when Z % 4 == 0; scalar tail otherwise)
__device__ __forceinline__ void sdf_pack_body(const SdfSource *__restrict__ src, int num_objects, int X, int Y, int Z,
                                              float *__restrict__ dst) {
    const int zq = (Z + 3) >> 2;
    const long long per_obj = (long long)X * Y * zq;
    const long long total = per_obj * num_objects;
    const bool vec = (Z & 3) == 0;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(g / per_obj);
        long long r = g - (long long)o * per_obj;
        const int z0 = (int)(r % zq) * 4;
        r /= zq;
        const int y = (int)(r % Y), x = (int)(r / Y);
        const SdfSource s = src[o];
        float v[4];
        const bool row_in = x < s.sx && y < s.sy;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int z = z0 + q;
            v[q] = (row_in && z < s.sz) ? sdf_source_at(s, x, y, z) : 1.0f;
        }
        float *out = dst + (((size_t)o * X + x) * Y + y) * Z + z0;
        if (vec) {
            *reinterpret_cast<float4 *>(out) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (z0 + q < Z) out[q] = v[q];
        }
    }
}

constexpr int POINT_TILE = 1024;

// One thread per voxel; the cloud streams through shared memory in tiles of POINT_TILE points.
__global__ void __launch_bounds__(256) point_sdf_kernel(const double *__restrict__ points, int num_points,
                                                        const double *__restrict__ gx, const double *__restrict__ gy,
                                                        const double *__restrict__ gz, int X, int Y, int Z,
                                                        float *__restrict__ out32, double *__restrict__ out64) {
    __shared__ double s_p[POINT_TILE * 3];
    const long long total = (long long)X * Y * Z;
    const long long base = (long long)blockIdx.x * blockDim.x;
    const long long v = base + threadIdx.x;
    const bool live = v < total;
    double px = 0.0, py = 0.0, pz = 0.0;
    if (live) {
        const int z = (int)(v % Z);
        const long long r = v / Z;
        px = __ldg(gx + (int)(r / Y)); py = __ldg(gy + (int)(r % Y)); pz = __ldg(gz + z);
    }
    double best = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    for (int t0 = 0; t0 < num_points; t0 += POINT_TILE) {
        const int cnt = min(POINT_TILE, num_points - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt * 3; k += blockDim.x) s_p[k] = __ldg(points + (size_t)t0 * 3 + k);
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int k = 0; k < cnt; ++k) {
                const double dx = __dsub_rn(px, s_p[3 * k]), dy = __dsub_rn(py, s_p[3 * k + 1]),
                             dz = __dsub_rn(pz, s_p[3 * k + 2]);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                best = fmin(best, d2);
            }
        }
    }
    if (live) {
        const double d = __dsqrt_rn(best);
        if (out64) out64[v] = d;
        if (out32) out32[v] = (float)d;
    }
}

}  // namespace omgb
