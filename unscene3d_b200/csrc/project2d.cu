// 2D -> 3D feature lifting: image features are carried along camera rays into the voxels they first hit.
//
// Replaces project_features_cuda.project_features_cuda (/root/reference/utils/cuda_utils/project_image_cuda_kernel.cu:24-64 ray
// march, :113-146 kernel, :183-256 host; caller utils/cuda_utils/raycast_image.py:18-77, used per image by
// pseudo_masks/unscene3d_pseudo_main.py:287-330 to lift DINO features before the NCut step).
//
// The reference runs one thread per (pixel, view) that marches its ray AND then loops over the feature channels with one scattered
// float atomic per channel (384 for DINO ViT-S/8): neighbouring threads write to unrelated voxels, so every atomic is its own
// memory transaction.  Here the work is split:
//   k_march    one thread per pixel: the ray march only, with the reference's arithmetic (same expressions in the same order, so
//              the float roundings — and with them the first occupied cell — agree), result = voxel index per pixel + hit counts;
//   k_scatter  one WARP per hit pixel: lanes stride the channels, a pixel's feature row is read coalesced and its reds land on
//              consecutive addresses of the voxel's row.
// Quirk kept: occupancy value 0 means "empty", so voxel 0 can never be hit (the reference stores voxel indices in the grid, :43-44).
#include "common.cuh"

namespace us3d {
namespace p2d {

struct Params {
    int B, V, H, W, C, Z, Y, X;
    float depth_min, depth_max, ray_inc;
};

__device__ __forceinline__ int sgn(float v) { return (0.0f < v) - (v < 0.0f); }

__global__ void __launch_bounds__(256) k_march(Params p, const long long *__restrict__ occ, const float *__restrict__ view_inv,
                                               const float *__restrict__ intr, int *__restrict__ hit, int *__restrict__ counts) {
    const long long total = (long long)p.B * p.V * p.H * p.W;
    const long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const int x = (int)(pix % p.W), y = (int)((pix / p.W) % p.H), view = (int)((pix / ((long long)p.W * p.H)) % p.V),
              batch = (int)(pix / ((long long)p.W * p.H * p.V));
    const float *m = view_inv + (size_t)(batch * p.V + view) * 16;
    const float fx = intr[batch * 4 + 0], fy = intr[batch * 4 + 1], mx = intr[batch * 4 + 2], my = intr[batch * 4 + 3];
    // kinectProjToCamera(depthMin, depthMax, mx, my, fx, fy, x, y, 1.0f) (cudaUtil.h:101-117), then normalize (cutil_math.h:1207)
    const float depth = 1.0f * (p.depth_max - p.depth_min) + p.depth_min;
    const float cx = ((float)(unsigned)x - mx) / fx, cy = ((float)(unsigned)y - my) / fy;
    float3 cam = make_float3(depth * cx, depth * cy, depth);
    {
        const float inv = rsqrtf(cam.x * cam.x + cam.y * cam.y + cam.z * cam.z);
        cam = make_float3(cam.x * inv, cam.y * inv, cam.z * inv);
    }
    // float4x4 * float3(0, 0, 0) (implicit w = 1) and float4x4 * float4(camDir, 0) (cuda_SimpleMatrixUtil.h:888-907)
    const float3 origin = make_float3(m[0] * 0.0f + m[1] * 0.0f + m[2] * 0.0f + m[3] * 1.0f, m[4] * 0.0f + m[5] * 0.0f + m[6] * 0.0f + m[7] * 1.0f,
                                      m[8] * 0.0f + m[9] * 0.0f + m[10] * 0.0f + m[11] * 1.0f);
    float3 dir = make_float3(m[0] * cam.x + m[1] * cam.y + m[2] * cam.z + m[3] * 0.0f, m[4] * cam.x + m[5] * cam.y + m[6] * cam.z + m[7] * 0.0f,
                             m[8] * cam.x + m[9] * cam.y + m[10] * cam.z + m[11] * 0.0f);
    {
        const float inv = rsqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
        dir = make_float3(dir.x * inv, dir.y * inv, dir.z * inv);
    }
    const float depth_to_ray = 1.0f / cam.z;
    float ray = depth_to_ray * p.depth_min;
    const float ray_end = depth_to_ray * p.depth_max;
    int found = 0;
#pragma unroll 1
    while (ray < ray_end) {
        const float3 w = make_float3(origin.x + ray * dir.x, origin.y + ray * dir.y, origin.z + ray * dir.z);
        const int px = (int)(w.x + (float)sgn(w.x) * 0.5f), py = (int)(w.y + (float)sgn(w.y) * 0.5f), pz = (int)(w.z + (float)sgn(w.z) * 0.5f);
        if (px >= 0 && py >= 0 && pz >= 0 && px < p.X && py < p.Y && pz < p.Z) {
            const int v = (int)occ[(size_t)batch * p.Z * p.Y * p.X + (size_t)pz * p.Y * p.X + (size_t)py * p.X + px];
            if (v != 0) {
                found = v;
                break;
            }
        }
        ray += p.ray_inc;
    }
    hit[pix] = found;
    if (found != 0) atomicAdd(&counts[found], 1);
}

template <typename T>
__global__ void __launch_bounds__(256) k_scatter(long long n_pix, int C, const int *__restrict__ hit, const T *__restrict__ feats,
                                                 T *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long pix = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); pix < n_pix; pix += warps) {
        const int v = hit[pix];
        if (v == 0) continue;
        const T *src = feats + (size_t)pix * C;
        T *dst = out + (size_t)v * C;
        for (int c = lane; c < C; c += 32) {
            if (sizeof(T) == sizeof(float) && std::is_floating_point<T>::value)
                atomicAdd(reinterpret_cast<float *>(dst + c), (float)src[c]);
            else
                atomicMax(reinterpret_cast<int *>(dst + c), (int)src[c]);
        }
    }
}

}  // namespace p2d
}  // namespace us3d

using namespace us3d;

extern "C" {

/* feats [B, V, H, W, C] (float: features, summed; int32 when pred_mode: labels, maximum), occ int64 [B, Z, Y, X] (0 = empty, else
 * voxel index), view_inv float [B, V, 4, 4] (camera-to-grid, in voxel units), intr float [B, 4] = (fx, fy, mx, my);
 * hit: int32 scratch [B * V * H * W] (receives the voxel index every pixel's ray hits first, 0 = none);
 * counts int32 [n_vox] and out [n_vox, C] are ACCUMULATED into (the caller zero-fills / presets them, as the reference's Python does). */
int us3d_project_features_2d3d(const void *feats, const long long *occ, const float *view_inv, const float *intr, int B, int V, int H, int W,
                               int C, int Z, int Y, int X, float depth_min, float depth_max, float ray_inc, int pred_mode, int *hit,
                               int *counts, void *out, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(B > 0 && V > 0 && H > 0 && W > 0 && C > 0 && Z > 0 && Y > 0 && X > 0, "project_features_2d3d: bad sizes");
    US3D_CHECK_ARG(ray_inc > 0.0f, "project_features_2d3d: the ray increment must be positive");
    p2d::Params p{B, V, H, W, C, Z, Y, X, depth_min, depth_max, ray_inc};
    const long long n_pix = (long long)B * V * H * W;
    p2d::k_march<<<(unsigned)((n_pix + 255) / 256), 256, 0, st>>>(p, occ, view_inv, intr, hit, counts);
    US3D_LAUNCH_CHECK();
    long long blocks = (n_pix + 7) / 8;
    if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
    if (pred_mode)
        p2d::k_scatter<int><<<(int)blocks, 256, 0, st>>>(n_pix, C, hit, (const int *)feats, (int *)out);
    else
        p2d::k_scatter<float><<<(int)blocks, 256, 0, st>>>(n_pix, C, hit, (const float *)feats, (float *)out);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
