// Tri-plane projection of sparse voxel predictions / targets for the noise-robust loss (custom_cuda_utils).
//
// Replaces project_sparse_voxels_to_planes[_backward] of /root/reference/utils/cuda_utils/cuda_utils_kernel.cu:371-433, 496-556
// (called from models/noise_robust_loss.py:28-31, 67-69).  Semantics kept, including the quirks: the planes are sized by the
// MAXIMUM centred coordinate (noise_robust_loss.py:84), so voxels ON the maximum of an axis fall outside and are skipped
// (kernel :392); the backward averages the three plane gradients over the ones that are non-zero (:539-548).
//
// The reference runs one thread per voxel looping over the instances (6 scattered float atomics per iteration, uncoalesced:
// neighbouring threads are `inst` floats apart).  Here one thread handles one (voxel, instance) element: a warp reads 32
// consecutive floats of a prediction row and its reds land on consecutive addresses of three plane rows.
#include "common.cuh"

namespace us3d {
namespace proj {

__device__ __forceinline__ bool cell(const int32_t *__restrict__ c, int v, int xd, int yd, int zd, int &x, int &y, int &z) {
    x = c[v * 4 + 1];
    y = c[v * 4 + 2];
    z = c[v * 4 + 3];
    return !(x >= xd || y >= yd || z >= zd || x < 0 || y < 0 || z < 0);
}

__global__ void __launch_bounds__(256) k_count(const int32_t *__restrict__ coords, int n, int xd, int yd, int zd,
                                               int *__restrict__ nxy, int *__restrict__ nxz, int *__restrict__ nyz) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    int x, y, z;
    if (v >= n || !cell(coords, v, xd, yd, zd, x, y, z)) return;
    atomicAdd(&nxy[x * yd + y], 1);
    atomicAdd(&nxz[x * zd + z], 1);
    atomicAdd(&nyz[y * zd + z], 1);
}

__global__ void __launch_bounds__(256) k_project(const int32_t *__restrict__ coords, const float *__restrict__ pred,
                                                 const float *__restrict__ tgt, int n, int inst, int xd, int yd, int zd,
                                                 float *__restrict__ pxy, float *__restrict__ pxz, float *__restrict__ pyz,
                                                 float *__restrict__ txy, float *__restrict__ txz, float *__restrict__ tyz) {
    const long long total = (long long)n * inst;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(e / inst), i = (int)(e - (long long)v * inst);
        int x, y, z;
        if (!cell(coords, v, xd, yd, zd, x, y, z)) continue;
        const float p = pred[e], t = tgt[e];
        const size_t a = (size_t)(x * yd + y) * inst + i, b = (size_t)(x * zd + z) * inst + i, c = (size_t)(y * zd + z) * inst + i;
        atomicAdd(pxy + a, p);
        atomicAdd(pxz + b, p);
        atomicAdd(pyz + c, p);
        atomicAdd(txy + a, t);
        atomicAdd(txz + b, t);
        atomicAdd(tyz + c, t);
    }
}

__global__ void __launch_bounds__(256) k_project_bwd(const int32_t *__restrict__ coords, int n, int inst, int xd, int yd, int zd,
                                                     const float *__restrict__ gxy, const float *__restrict__ gxz,
                                                     const float *__restrict__ gyz, float *__restrict__ grad) {
    const long long total = (long long)n * inst;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(e / inst), i = (int)(e - (long long)v * inst);
        int x, y, z;
        if (!cell(coords, v, xd, yd, zd, x, y, z)) continue;  // the reference leaves these entries of s_grads untouched
        const float a = gxy[(size_t)(x * yd + y) * inst + i], b = gxz[(size_t)(x * zd + z) * inst + i],
                    c = gyz[(size_t)(y * zd + z) * inst + i];
        const int cnt = (a != 0.0f) + (b != 0.0f) + (c != 0.0f);
        grad[e] = cnt > 0 ? (a + b + c) / cnt : 0.0f;
    }
}

}  // namespace proj
}  // namespace us3d

using namespace us3d;

extern "C" {

int us3d_project_voxels_to_planes(const int32_t *coords, const float *pred, const float *tgt, int n, int inst, int x_dim, int y_dim,
                                  int z_dim, float *pred_xy, float *pred_xz, float *pred_yz, float *tgt_xy, float *tgt_xz,
                                  float *tgt_yz, int32_t *num_xy, int32_t *num_xz, int32_t *num_yz, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n >= 0 && inst > 0 && x_dim >= 0 && y_dim >= 0 && z_dim >= 0, "project_voxels_to_planes: bad sizes");
    if (n == 0 || x_dim == 0 || y_dim == 0 || z_dim == 0) return 0;
    proj::k_count<<<ceil_div(n, 256), 256, 0, st>>>(coords, n, x_dim, y_dim, z_dim, num_xy, num_xz, num_yz);
    US3D_LAUNCH_CHECK();
    long long blocks = ((long long)n * inst + 255) / 256;
    if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
    proj::k_project<<<(int)blocks, 256, 0, st>>>(coords, pred, tgt, n, inst, x_dim, y_dim, z_dim, pred_xy, pred_xz, pred_yz, tgt_xy,
                                                 tgt_xz, tgt_yz);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_project_voxels_to_planes_bwd(const int32_t *coords, int n, int inst, int x_dim, int y_dim, int z_dim, const float *grad_xy,
                                      const float *grad_xz, const float *grad_yz, float *grad, void *stream_) {
    US3D_CHECK_ARG(n >= 0 && inst > 0, "project_voxels_to_planes_bwd: bad sizes");
    if (n == 0 || x_dim == 0 || y_dim == 0 || z_dim == 0) return 0;
    long long blocks = ((long long)n * inst + 255) / 256;
    if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
    proj::k_project_bwd<<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>(coords, n, inst, x_dim, y_dim, z_dim, grad_xy, grad_xz, grad_yz,
                                                                        grad);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
