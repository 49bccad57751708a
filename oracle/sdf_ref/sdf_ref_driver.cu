// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/ load the library built from this file.
//
// Pins the operator oracle (oracle/sdf_loss_ref.c) and the product operator (omgb_sdf_loss) to the REFERENCE'S OWN
// SOURCE: the translation unit below is /root/reference/layers/sdf_matching_loss_kernel.cu, included where it lies
// (nothing of it is copied into this repository), compiled by nvcc over the stand-in headers of oracle/sdf_ref/shim
// (ATen / Eigen / Sophus / thrust are absent from this image; see each shim header for what it restates).
//
//   sdfref_value_interp / sdfref_grad_interp   host calls of the reference's __device__ __host__ helpers
//       getValueInterpolated (kernel.cu:37-64) and getGradientInterpolated (kernel.cu:67-86) -- runs on any CPU.
//       The host pass is compiled with -mfma -ffp-contract=fast so that lerp (kernel.cu:15-18) rounds as the device
//       build's FFMA does; sdfref_forward_device below proves on the B200 that this is what nvcc emits.
//   sdfref_forward_device                     the reference's own host function sdf_loss_cuda_forward
//       (kernel.cu:204-262) with its own kernels SDFdistanceForward / sum_gradients, on the current CUDA device.
#include REF_KERNEL_CU

extern "C" {

// pgrid [N,3] grid coordinates; grid [d0,d1,d2]; out [N]
void sdfref_value_interp(const float *pgrid, int n, const float *grid, int d0, int d1, int d2, float *out) {
    const int3 dim = make_int3(d0, d1, d2);
    for (int i = 0; i < n; ++i)
        out[i] = getValueInterpolated<float>(make_float3(pgrid[3 * i], pgrid[3 * i + 1], pgrid[3 * i + 2]), dim, grid);
}

// out [N,3]
void sdfref_grad_interp(const float *pgrid, int n, const float *grid, int d0, int d1, int d2, float delta, float *out) {
    const int3 dim = make_int3(d0, d1, d2);
    for (int i = 0; i < n; ++i) {
        const float3 g = getGradientInterpolated<float>(make_float3(pgrid[3 * i], pgrid[3 * i + 1], pgrid[3 * i + 2]),
                                                        dim, grid, delta);
        out[3 * i] = g.x; out[3 * i + 1] = g.y; out[3 * i + 2] = g.z;
    }
}

// All pointers are DEVICE pointers (fp32, contiguous), same contract as omg_cuda.sdf_loss_forward
// (layers/omg_layers.cpp:24-49).  Returns 0, or a cudaError_t.
int sdfref_forward_device(float *pose_init, float *sdf_grids, float *sdf_limits, float *points, float *epsilons,
                          float *padding_scales, float *clearances, float *disables, int num_points, int num_objects,
                          int d0, int d1, int d2, float *potentials, float *potential_grads, float *collides) {
    at::Tensor t_pose(pose_init, {num_objects, 4, 4}), t_grids(sdf_grids, {num_objects, d0, d1, d2}),
        t_lim(sdf_limits, {num_objects, 10}), t_pts(points, {num_points, 3}), t_eps(epsilons, {num_objects}),
        t_pad(padding_scales, {num_objects}), t_clr(clearances, {num_objects}), t_dis(disables, {num_objects});
    std::vector<at::Tensor> out = sdf_loss_cuda_forward(t_pose, t_grids, t_lim, t_pts, t_eps, t_pad, t_clr, t_dis);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const size_t n = (size_t)num_points;
    if ((e = cudaMemcpy(potentials, out[0].data<float>(), sizeof(float) * n, cudaMemcpyDeviceToDevice)) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpy(potential_grads, out[1].data<float>(), sizeof(float) * 3 * n, cudaMemcpyDeviceToDevice)) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpy(collides, out[2].data<float>(), sizeof(float) * n, cudaMemcpyDeviceToDevice)) != cudaSuccess) return (int)e;
    return (int)cudaDeviceSynchronize();
}

}  // extern "C"
