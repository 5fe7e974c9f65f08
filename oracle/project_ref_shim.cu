// extern "C" entry around the reference's own 2D -> 3D feature projection
// (/root/reference/utils/cuda_utils/project_image_cuda_kernel.cu:24-64 ray march, :113-146 kernel, :183-256 host function).
// Compiled together with that file into oracle/_ref/libproject_ref.so by oracle/build_ref.py.  TEST INFRASTRUCTURE: the checker
// for us3d_project_features_2d3d.
#include <torch/extension.h>
#include <cuda_runtime.h>

void project_features_cuda_forward(at::Tensor encoded_2d_features, at::Tensor occupancy_3D, at::Tensor viewMatrixInv, at::Tensor intrinsicParams,
                                   at::Tensor opts, at::Tensor mapping2dto3d_num, at::Tensor projected_features, at::Tensor pred_mode_t);

extern "C" int project_ref(const float *feats, const int64_t *occ, const float *views, const float *intr, int B, int V, int H, int W, int C,
                           int Z, int Y, int X, float depth_min, float depth_max, float ray_inc, int n_vox, int *counts, float *out) {
    auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA);
    auto i32 = torch::TensorOptions().dtype(torch::kInt32).device(torch::kCUDA);
    auto i64 = torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA);
    at::Tensor opts = torch::tensor({(float)W, (float)H, depth_min, depth_max, ray_inc});
    at::Tensor pred = torch::zeros({1}, torch::kBool);
    project_features_cuda_forward(torch::from_blob((void *)feats, {B, V, H, W, C}, f32), torch::from_blob((void *)occ, {B, Z, Y, X}, i64),
                                  torch::from_blob((void *)views, {B, V, 4, 4}, f32), torch::from_blob((void *)intr, {B, 4}, f32), opts,
                                  torch::from_blob((void *)counts, {n_vox}, i32), torch::from_blob((void *)out, {n_vox, C}, f32), pred);
    return (int)cudaDeviceSynchronize();
}
