// SDF sampling device code shared by the raw operator kernel and the fused CHOMP kernels.
//
// Arithmetic contract (what "results identical to the reference" means for this operator): every fp32
// operation of layers/sdf_matching_loss_kernel.cu:97-181 is reproduced as a separately rounded IEEE op,
// with fused multiply-adds exactly where nvcc 12.9 fuses them when it compiles the reference's own source for
// sm_100a (oracle/sdf_ref builds that source; its SASS was read instruction by instruction and the result is
// checked bit for bit on the B200, tests/test_gpu_ref_operator.py): the lerp a + t*(b-a), the quaternion cross
// products, w*uv + p, the quaternion -> rotation matrix entries and the R^T*g accumulation.  The reference's double-precision steps
// ((p - 0.5) in double, 0.5*(f+ - f-)/delta in double, -v + 0.5*eps in double) are evaluated in fp32 in a
// way that is bit-identical to the double path (double rounding through binary64 is innocuous for +,-,/
// of binary32 operands since 53 >= 2*24+2; the one value where (int)(p - 0.5) would differ is handled by
// comparing p against -0.5 directly).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace omgb {

// Per-object record prepared once per pose/parameter update (prep_objects_kernel).
struct __align__(16) ObjRec {
    float qw, qx, qy, qz;        // unit quaternion of the world->object rotation (Eigen matrix->quaternion)
    float tx, ty, tz, delta;     // translation; voxel size (sdf_limits[9])
    float r[9];                  // quaternion -> rotation matrix (what so3().matrix() returns)
    float minx, miny, minz;      // sdf_limits[0:3]
    float ex, ey, ez;            // sdf_limits[3:6] - sdf_limits[0:3]
    float fd0, fd1, fd2;         // (float)dims
    int d0, d1, d2;              // dims
    float eps, pad, clr, dis;    // epsilons, padding_scales, clearances, disables
    float inv2eps, inveps;       // 1/(2*eps), 1/eps  (kernel.cu:167-168)
    float lox, loy, loz;         // object-frame interval inside which a point's 8-tap cell is in bounds
    float hix, hiy, hiz;
    float cull_pad;              // conservative slack for the sphere cull
    float isx, isy, isz;         // 1 / grid spacing per axis (cull only)
    float alox, aloy, aloz;      // object-frame AABB of the region where a sample can be <= eps or < clearance
    float ahix, ahiy, ahiz;      // (from the lower-bound grid; empty when alo > ahi)
    int dil_off;                 // o * (bricks per object) into the lower-bound grid
    long long grid_offset;       // o * d0*d1*d2
    float ga[12];                // world -> APPROXIMATE grid coordinates, g = ga[4k..4k+2] . x + ga[4k+3] (cull/classify only)
    float wsx, wsy, wsz, wsr;    // WORLD-frame sphere enclosing the padded in-bounds box (first-level cull); wsr < 0: never cull
    long long quad_offset;       // o * QuadDesc::obj_stride (float4 units) into the bricked quad copy of the grids
};

// Optional second copy of the packed SDFs for the exact-evaluation path ("bricked quads", omgb_scene_set_sdf_layout):
// one float4 per voxel cell (x, y, z) = the four taps (x,y,z) (x,y,z+1) (x,y+1,z) (x,y+1,z+1), cells grouped in 8^3
// bricks of 8 KB, x fastest inside a brick.  A trilinear sample is then TWO aligned 128-bit loads (cells x0 and x0+1,
// usually one 32-byte sector) instead of eight scalar taps over four (x, y) rows; same tap values, same lerp order,
// so results are bit-identical.  4x the bytes of the reference layout, which stays in place for the raw operator.
struct QuadDesc {
    const float4 *data;          // null: read the reference [O,X,Y,Z] layout
    long long obj_stride;        // float4s per object = nb0 * nb1 * nb2 * 512
    int nb1, nb2;                // bricks along y and z
};

__host__ __device__ __forceinline__ long long quad_index(int x, int y, int z, int nb1, int nb2) {
    const long long brick = ((long long)(x >> 3) * nb1 + (y >> 3)) * nb2 + (z >> 3);
    return (brick << 9) + ((((y & 7) << 3) + (z & 7)) << 3) + (x & 7);
}

__device__ __forceinline__ float lerp_ref(float a, float b, float t) {   // kernel.cu:15-18
    return __fmaf_rn(t, __fsub_rn(b, a), a);
}

// (int)(p - 0.5) and (float)((p - 0.5) - x0) of kernel.cu:39-41, negative cell index => out of bounds.
__device__ __forceinline__ void cell_of(float p, int &c0, float &f) {
    const float t = __fsub_rn(p, 0.5f);
    if (p >= 0.5f) {
        c0 = __float2int_rz(t);
        f = __fsub_rn(t, (float)c0);
    } else {
        c0 = (p > -0.5f) ? 0 : -1;
        f = t;
    }
}

// kernel.cu:37-64.  Returns 1.0f when any of the 8 taps is out of bounds.  q != null: the object's bricked quads.
__device__ __forceinline__ float value_interp(const float *__restrict__ g, const float4 *__restrict__ q, int nb1,
                                              int nb2, int d0, int d1, int d2, float px, float py, float pz,
                                              bool &inb) {
    int x0, y0, z0;
    float fx, fy, fz;
    cell_of(px, x0, fx);
    cell_of(py, y0, fy);
    cell_of(pz, z0, fz);
    inb = (x0 >= 0) & (x0 + 1 < d0) & (y0 >= 0) & (y0 + 1 < d1) & (z0 >= 0) & (z0 + 1 < d2);
    if (!inb) return 1.0f;
    float v000, v001, v010, v011, v100, v101, v110, v111;
    if (q) {   // two LDG.E.128
        const float4 a = __ldg(q + quad_index(x0, y0, z0, nb1, nb2));
        const float4 c = __ldg(q + quad_index(x0 + 1, y0, z0, nb1, nb2));
        v000 = a.x; v001 = a.y; v010 = a.z; v011 = a.w;
        v100 = c.x; v101 = c.y; v110 = c.z; v111 = c.w;
    } else {
        const float *b = g + ((size_t)x0 * d1 + y0) * d2 + z0;
        const size_t sx = (size_t)d1 * d2;
        v000 = __ldg(b); v001 = __ldg(b + 1);
        v010 = __ldg(b + d2); v011 = __ldg(b + d2 + 1);
        v100 = __ldg(b + sx); v101 = __ldg(b + sx + 1);
        v110 = __ldg(b + sx + d2); v111 = __ldg(b + sx + d2 + 1);
    }
    const float dx00 = lerp_ref(v000, v100, fx);
    const float dx01 = lerp_ref(v001, v101, fx);
    const float dx10 = lerp_ref(v010, v110, fx);
    const float dx11 = lerp_ref(v011, v111, fx);
    const float dxy0 = lerp_ref(dx00, dx10, fy);
    const float dxy1 = lerp_ref(dx01, dx11, fy);
    return lerp_ref(dxy0, dxy1, fz);
}

// world point -> grid coordinates of object `o` (kernel.cu:125-142; so3.hpp:298-300 quaternion rotation).
__device__ __forceinline__ void to_grid(const ObjRec &o, float x, float y, float z, float &px, float &py,
                                        float &pz) {
    float ux = __fmaf_rn(o.qy, z, -__fmul_rn(o.qz, y));
    float uy = __fmaf_rn(o.qz, x, -__fmul_rn(o.qx, z));
    float uz = __fmaf_rn(o.qx, y, -__fmul_rn(o.qy, x));
    ux = __fadd_rn(ux, ux); uy = __fadd_rn(uy, uy); uz = __fadd_rn(uz, uz);
    const float cx = __fmaf_rn(o.qy, uz, -__fmul_rn(o.qz, uy));
    const float cy = __fmaf_rn(o.qz, ux, -__fmul_rn(o.qx, uz));
    const float cz = __fmaf_rn(o.qx, uy, -__fmul_rn(o.qy, ux));
    const float wx = __fadd_rn(__fadd_rn(__fmaf_rn(o.qw, ux, x), cx), o.tx);
    const float wy = __fadd_rn(__fadd_rn(__fmaf_rn(o.qw, uy, y), cy), o.ty);
    const float wz = __fadd_rn(__fadd_rn(__fmaf_rn(o.qw, uz, z), cz), o.tz);
    px = __fmul_rn(__fdiv_rn(__fsub_rn(wx, o.minx), o.ex), o.fd0);
    py = __fmul_rn(__fdiv_rn(__fsub_rn(wy, o.miny), o.ey), o.fd1);
    pz = __fmul_rn(__fdiv_rn(__fsub_rn(wz, o.minz), o.ez), o.fd2);
}

// Value-only evaluation of one (point, object) pair: potential + collide flag (kernel.cu:147-171 without
// the gradient).  Returns true if the 8-tap cell is in bounds.
__device__ __forceinline__ bool pair_potential(const ObjRec &o, const float *__restrict__ grids, const QuadDesc &qd,
                                               float x, float y, float z, float &pot, float &col) {
    float px, py, pz;
    to_grid(o, x, y, z, px, py, pz);
    bool inb;
    const float v = value_interp(grids + o.grid_offset, qd.data ? qd.data + o.quad_offset : nullptr, qd.nb1, qd.nb2,
                                 o.d0, o.d1, o.d2, px, py, pz, inb);
    col = (v < o.clr) ? 1.0f : 0.0f;
    if (v <= 0.0f) {
        pot = __fadd_rn(-v, __fmul_rn(0.5f, o.eps));
    } else if (v <= o.eps) {
        const float d = __fsub_rn(v, o.eps);
        pot = __fmul_rn(__fmul_rn(__fmul_rn(o.inv2eps, d), d), o.pad);
    } else {
        pot = 0.0f;
    }
    return inb;
}

// Potential, world-frame potential gradient and collide flag of one pair from its seven trilinear samples
// (value, then +x +y +z -x -y -z shifted by one voxel): kernel.cu:150-180.
__device__ __forceinline__ void finish_pair(const ObjRec &o, float v, float fpx, float fpy, float fpz, float fmx,
                                            float fmy, float fmz, float &pot, float &gx, float &gy, float &gz,
                                            float &col) {
    col = (v < o.clr) ? 1.0f : 0.0f;
    pot = 0.0f; gx = gy = gz = 0.0f;
    if (!(v <= o.eps)) return;        // value > eps (including the OOB 1.0 when eps < 1): kernel.cu:172-173
    const float dgx = __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(fpx, fmx)), o.delta);   // kernel.cu:82-84
    const float dgy = __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(fpy, fmy)), o.delta);
    const float dgz = __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(fpz, fmz)), o.delta);
    float vx, vy, vz;
    if (v <= 0.0f) {                                                   // kernel.cu:158-164
        pot = __fadd_rn(-v, __fmul_rn(0.5f, o.eps));
        vx = -dgx; vy = -dgy; vz = -dgz;
    } else {                                                           // kernel.cu:165-171
        const float d = __fsub_rn(v, o.eps);
        pot = __fmul_rn(__fmul_rn(__fmul_rn(o.inv2eps, d), d), o.pad);
        vx = __fmul_rn(__fmul_rn(__fmul_rn(o.inveps, dgx), d), o.pad);
        vy = __fmul_rn(__fmul_rn(__fmul_rn(o.inveps, dgy), d), o.pad);
        vz = __fmul_rn(__fmul_rn(__fmul_rn(o.inveps, dgz), d), o.pad);
    }
    // rotationMatrix.transpose() * vgrad   (kernel.cu:176): Eigen's 3-term tree sum c0 + (c1 + c2) as nvcc fuses it
    gx = __fmaf_rn(o.r[0], vx, __fmaf_rn(o.r[3], vy, __fmul_rn(o.r[6], vz)));
    gy = __fmaf_rn(o.r[1], vx, __fmaf_rn(o.r[4], vy, __fmul_rn(o.r[7], vz)));
    gz = __fmaf_rn(o.r[2], vx, __fmaf_rn(o.r[5], vy, __fmul_rn(o.r[8], vz)));
}

// Full evaluation by one thread: potential, world-frame potential gradient, collide flag (kernel.cu:147-180).
__device__ __forceinline__ bool pair_full(const ObjRec &o, const float *__restrict__ grids, const QuadDesc &qd,
                                          float x, float y, float z, float &pot, float &gx, float &gy, float &gz,
                                          float &col) {
    float px, py, pz;
    to_grid(o, x, y, z, px, py, pz);
    const float *g = grids + o.grid_offset;
    const float4 *q = qd.data ? qd.data + o.quad_offset : nullptr;
    const int n1 = qd.nb1, n2 = qd.nb2;
    bool inb, dummy;
    const float v = value_interp(g, q, n1, n2, o.d0, o.d1, o.d2, px, py, pz, inb);
    float fpx = 0.f, fpy = 0.f, fpz = 0.f, fmx = 0.f, fmy = 0.f, fmz = 0.f;
    if (v <= o.eps) {   // kernel.cu:67-86: six re-interpolations (their result is unused when value > eps)
        fpx = value_interp(g, q, n1, n2, o.d0, o.d1, o.d2, __fadd_rn(px, 1.0f), py, pz, dummy);
        fpy = value_interp(g, q, n1, n2, o.d0, o.d1, o.d2, px, __fadd_rn(py, 1.0f), pz, dummy);
        fpz = value_interp(g, q, n1, n2, o.d0, o.d1, o.d2, px, py, __fadd_rn(pz, 1.0f), dummy);
        fmx = value_interp(g, q, n1, n2, o.d0, o.d1, o.d2, __fsub_rn(px, 1.0f), py, pz, dummy);
        fmy = value_interp(g, q, n1, n2, o.d0, o.d1, o.d2, px, __fsub_rn(py, 1.0f), pz, dummy);
        fmz = value_interp(g, q, n1, n2, o.d0, o.d1, o.d2, px, py, __fsub_rn(pz, 1.0f), dummy);
    }
    finish_pair(o, v, fpx, fpy, fpz, fmx, fmy, fmz, pot, gx, gy, gz, col);
    return inb;
}

// The same evaluation spread over a group of G lanes (G = 1, 2, 4 or 8; gm = the group's lane mask, l = lane
// within the group): lane l takes samples l, l+G, ... of the seven, the values are exchanged with shuffles and
// every lane finishes redundantly.  G == 1 evaluates the six gradient samples only when value <= eps.
template <int G>
__device__ __forceinline__ void pair_full_group(const ObjRec &o, const float *__restrict__ grids, const QuadDesc &qd,
                                                unsigned gm, int l, float x, float y, float z, float &pot, float &gx,
                                                float &gy, float &gz, float &col) {
    if (G == 1) {
        pair_full(o, grids, qd, x, y, z, pot, gx, gy, gz, col);
        return;
    }
    float px, py, pz;
    to_grid(o, x, y, z, px, py, pz);
    constexpr int Q = (7 + G - 1) / G;
    float v[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int k = l + q * G;
        v[q] = 0.0f;
        if (k < 7) {
            // p + 1 and p - 1 are single roundings either way (float3 operator+/- of kernel.cu:20-28)
            const float sx = (k == 1) ? 1.0f : ((k == 4) ? -1.0f : 0.0f);
            const float sy = (k == 2) ? 1.0f : ((k == 5) ? -1.0f : 0.0f);
            const float sz = (k == 3) ? 1.0f : ((k == 6) ? -1.0f : 0.0f);
            bool inb;
            v[q] = value_interp(grids + o.grid_offset, qd.data ? qd.data + o.quad_offset : nullptr, qd.nb1, qd.nb2, o.d0,
                                o.d1, o.d2, __fadd_rn(px, sx), __fadd_rn(py, sy), __fadd_rn(pz, sz), inb);
        }
    }
    float f[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) f[k] = __shfl_sync(gm, v[k / G], k % G, G);
    finish_pair(o, f[0], f[1], f[2], f[3], f[4], f[5], f[6], pot, gx, gy, gz, col);
}

}  // namespace omgb
