// TEST INFRASTRUCTURE (oracle/sdf_ref): stand-in for <ATen/ATen.h>, written for this repo, so that the reference's
// layers/sdf_matching_loss_kernel.cu compiles WHERE IT LIES without libtorch headers.  It provides exactly what that
// file uses of at::Tensor (kernel.cu:204-262): size(), options(), data<float>() and at::zeros() -- a reference-counted
// device allocation.  No arithmetic lives here.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <memory>
#include <vector>

namespace at {

struct TensorOptions {};

class Tensor {
  public:
    Tensor() {}
    // borrowed device memory (the driver wraps the caller's buffers)
    Tensor(float *borrowed, std::initializer_list<long long> sizes) : ptr_(borrowed), sizes_(sizes) {}
    long long size(int d) const { return sizes_[(size_t)d]; }
    TensorOptions options() const { return TensorOptions(); }
    template <typename T> T *data() const { return reinterpret_cast<T *>(ptr_); }
    long long numel() const {
        long long n = 1;
        for (long long s : sizes_) n *= s;
        return n;
    }

  private:
    friend Tensor zeros(std::initializer_list<long long>, const TensorOptions &);
    float *ptr_ = nullptr;
    std::vector<long long> sizes_;
    std::shared_ptr<void> own_;
};

inline Tensor zeros(std::initializer_list<long long> sizes, const TensorOptions &) {
    Tensor t;
    t.sizes_ = sizes;
    const size_t bytes = sizeof(float) * (size_t)(t.numel() > 0 ? t.numel() : 1);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess) {
        fprintf(stderr, "oracle/sdf_ref shim: device allocation of %zu bytes failed\n", bytes);
        abort();
    }
    t.ptr_ = reinterpret_cast<float *>(p);
    t.own_ = std::shared_ptr<void>(p, [](void *q) { cudaFree(q); });
    return t;
}

}  // namespace at
