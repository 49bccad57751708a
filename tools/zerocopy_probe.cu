// Diagnostic (run on the GPU box): how fast do SMs read mapped pinned HOST memory, by access pattern?
// One CTA per 2160-byte record (a 30-waypoint trajectory), 1024 records -- the staging pattern of the zero-copy
// entry point -- with (a) 8-byte loads, (b) 16-byte loads, (c) one cp.async.bulk per record, (d) 4-byte loads;
// plus the copy engine (cudaMemcpyAsync) for the same bytes.  nvcc -arch=sm_100a -O3 -o /tmp/zc tools/zerocopy_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(320) read_kernel(const double *src, int rec_doubles, double *sink) {
    extern __shared__ __align__(16) double s[];
    __shared__ __align__(8) unsigned long long bar;
    const double *g = src + (size_t)blockIdx.x * rec_doubles;
    const int tid = threadIdx.x;
    double acc = 0.0;
    if (MODE == 0) {
        for (int k = tid; k < rec_doubles; k += blockDim.x) s[k] = __ldcg(g + k);
    } else if (MODE == 1) {
        const double2 *g2 = reinterpret_cast<const double2 *>(g);
        double2 *s2 = reinterpret_cast<double2 *>(s);
        for (int k = tid; k < rec_doubles / 2; k += blockDim.x) s2[k] = __ldcg(g2 + k);
    } else if (MODE == 3) {
        const float *g1 = reinterpret_cast<const float *>(g);
        float *s1 = reinterpret_cast<float *>(s);
        for (int k = tid; k < rec_doubles * 2; k += blockDim.x) s1[k] = __ldcg(g1 + k);
    } else {
        const unsigned b = smem_u32(&bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(rec_doubles * 8) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(s)), "l"(g), "r"(rec_doubles * 8), "r"(b) : "memory");
        }
        __syncthreads();
        unsigned ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(b), "r"(0u) : "memory");
    }
    __syncthreads();
    for (int k = tid; k < rec_doubles; k += blockDim.x) acc += s[k];
    if (acc == 123.456) sink[0] = acc;
}

template <int MODE>
__global__ void __launch_bounds__(320) write_kernel(double *dst, int rec_doubles) {
    double *g = dst + (size_t)blockIdx.x * rec_doubles;
    if (MODE == 0) for (int k = threadIdx.x; k < rec_doubles; k += blockDim.x) g[k] = (double)k;
    else {
        double2 *g2 = reinterpret_cast<double2 *>(g);
        for (int k = threadIdx.x; k < rec_doubles / 2; k += blockDim.x) g2[k] = make_double2(k, k);
    }
}

int main() {
    const int recs = 1024, rd = 270;   // 2160 B per record
    const size_t bytes = (size_t)recs * rd * 8;
    double *h, *d, *sink;
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocMapped));
    for (size_t i = 0; i < bytes / 8; ++i) h[i] = (double)i;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&sink, 8));
    double *hd; CK(cudaHostGetDevicePointer(&hd, h, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t smem = rd * 8;
    float ms;
#define TIME(label, launch)                                                                     \
    do {                                                                                        \
        for (int w = 0; w < 3; ++w) { launch; }                                                 \
        CK(cudaDeviceSynchronize());                                                            \
        float best = 1e9f, sum = 0;                                                             \
        for (int r = 0; r < 10; ++r) {                                                          \
            CK(cudaEventRecord(e0)); launch; CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
            CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; sum += ms;     \
        }                                                                                       \
        printf("%-54s best %.4f ms  mean %.4f ms  -> %.1f GB/s (best)\n", label, best, sum / 10, bytes / (best * 1e6)); \
    } while (0)
    TIME("host->SM   8-byte loads, CTA per 2160-B record", (read_kernel<0><<<recs, 320, smem>>>(hd, rd, sink)));
    TIME("host->SM  16-byte loads, CTA per 2160-B record", (read_kernel<1><<<recs, 320, smem>>>(hd, rd, sink)));
    TIME("host->SM   4-byte loads, CTA per 2160-B record", (read_kernel<3><<<recs, 320, smem>>>(hd, rd, sink)));
    TIME("host->SM  cp.async.bulk per 2160-B record", (read_kernel<2><<<recs, 320, smem>>>(hd, rd, sink)));
    TIME("device->SM 8-byte loads (reference)", (read_kernel<0><<<recs, 320, smem>>>(d, rd, sink)));
    TIME("SM->host   8-byte stores", (write_kernel<0><<<recs, 320>>>(hd, rd)));
    TIME("SM->host  16-byte stores", (write_kernel<1><<<recs, 320>>>(hd, rd)));
    TIME("copy engine H2D (cudaMemcpyAsync)", (cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)));
    TIME("copy engine D2H (cudaMemcpyAsync)", (cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost)));
    // only 148 CTAs looping over records (persistent): fewer requesters in flight
    return 0;
}
