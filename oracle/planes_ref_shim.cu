// extern "C" entries around the reference's own tri-plane projection functions
// (/root/reference/utils/cuda_utils/cuda_utils_kernel.cu:436-493 forward, :559-600 backward; kernels :371-433, 496-556).
// Compiled together with that file into oracle/_ref/libplanes_ref.so by oracle/build_ref.py; raw device pointers are wrapped
// into at::Tensor views with torch::from_blob.  TEST INFRASTRUCTURE: the checker for us3d_project_voxels_to_planes[_bwd].
#include <torch/extension.h>
#include <cuda_runtime.h>

void project_sparse_voxels_to_planes(at::Tensor s_coords, at::Tensor s_predictions, at::Tensor s_targets, at::Tensor xy_pred_projections,
                                     at::Tensor xz_pred_projections, at::Tensor yz_pred_projections, at::Tensor xy_target_projections,
                                     at::Tensor xz_target_projections, at::Tensor yz_target_projections, at::Tensor xy_projection_nums,
                                     at::Tensor xz_projection_nums, at::Tensor yz_projection_nums);
void project_sparse_voxels_to_planes_backward(at::Tensor s_coords, at::Tensor s_grads, at::Tensor xy_grads, at::Tensor xz_grads,
                                              at::Tensor yz_grads, at::Tensor xy_nums, at::Tensor xz_nums, at::Tensor yz_nums);

static at::Tensor f32(const void *p, std::vector<int64_t> shape) {
    return torch::from_blob(const_cast<void *>(p), shape, torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA));
}
static at::Tensor i32(const void *p, std::vector<int64_t> shape) {
    return torch::from_blob(const_cast<void *>(p), shape, torch::TensorOptions().dtype(torch::kInt32).device(torch::kCUDA));
}

extern "C" int planes_ref_fwd(const int *coords, const float *pred, const float *tgt, int n, int inst, int xd, int yd, int zd, float *pxy,
                              float *pxz, float *pyz, float *txy, float *txz, float *tyz, int *nxy, int *nxz, int *nyz) {
    project_sparse_voxels_to_planes(i32(coords, {n, 4}), f32(pred, {n, inst}), f32(tgt, {n, inst}), f32(pxy, {xd, yd, inst}),
                                    f32(pxz, {xd, zd, inst}), f32(pyz, {yd, zd, inst}), f32(txy, {xd, yd, inst}), f32(txz, {xd, zd, inst}),
                                    f32(tyz, {yd, zd, inst}), i32(nxy, {xd, yd}), i32(nxz, {xd, zd}), i32(nyz, {yd, zd}));
    return (int)cudaDeviceSynchronize();
}

extern "C" int planes_ref_bwd(const int *coords, float *grads, int n, int inst, int xd, int yd, int zd, const float *gxy, const float *gxz,
                              const float *gyz, const int *nxy, const int *nxz, const int *nyz) {
    project_sparse_voxels_to_planes_backward(i32(coords, {n, 4}), f32(grads, {n, inst}), f32(gxy, {xd, yd, inst}), f32(gxz, {xd, zd, inst}),
                                             f32(gyz, {yd, zd, inst}), i32(nxy, {xd, yd}), i32(nxz, {xd, zd}), i32(nyz, {yd, zd}));
    return (int)cudaDeviceSynchronize();
}
