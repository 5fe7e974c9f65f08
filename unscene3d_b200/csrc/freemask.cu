// FreeMask-style pseudo-mask variant: the segment branch of the scene loop in the reference's
// pseudo_masks/freemask_main.py:203-417 — all-pairs cosine soft masks between segment features, hard threshold,
// separation of non-connected blobs, maskness ranking, extent filter and mask-IoU suppression (utils/pc_utils.py:724-757).
//
// The reference maps every candidate mask onto the points ([M, N] floats, :359-372) and runs the suppression as a double
// Python loop over those rows (O(M^2 N)).  Every candidate is a union of whole segments, so here nothing of size M x N
// exists: point counts, bounding boxes and pairwise intersections are computed at segment level with per-segment
// weights (number of points, bounding box) and are exactly the integers the reference gets at point level.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace us3d {
namespace fm {

// ||f_i|| in fp32 like torch's norm (fp64 accumulation, rounded once)
__global__ void __launch_bounds__(256) k_row_norm(const float *__restrict__ f, int s, int d, float *__restrict__ nrm) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= s) return;
    double acc = 0;
    for (int c = lane; c < d; c += 32) {
        double v = f[(size_t)row * d + c];
        acc += v * v;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) nrm[row] = (float)sqrt(acc);
}

// A[i, j] = < f_i / (||f_i|| + eps), f_j / (||f_j|| + eps) >   (utils/freemask_utils.py:12-15; rows = queries, columns = keys)
__global__ void __launch_bounds__(256) k_cosine(const float *__restrict__ f, const float *__restrict__ nrm, int s, int d,
                                                float *__restrict__ A) {
    __shared__ float ta[32][33], tb[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const float eps = 10e-10f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < d; c0 += 32) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int rr = ty + 8 * r, c = c0 + tx;
            ta[rr][tx] = (i0 + rr < s && c < d) ? f[(size_t)(i0 + rr) * d + c] / (nrm[i0 + rr] + eps) : 0.f;
            tb[rr][tx] = (j0 + rr < s && c < d) ? f[(size_t)(j0 + rr) * d + c] / (nrm[j0 + rr] + eps) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            float b = tb[tx][c];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(ta[ty + 8 * r][c], b, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int i = i0 + ty + 8 * r, j = j0 + tx;
        if (i < s && j < s) A[(size_t)i * s + j] = acc[r];
    }
}

// attn -= rowmin; attn /= rowmax + eps (:16-17); columns of all-zero keys <- 0 (freemask_main.py:266).  One warp per row.
__global__ void __launch_bounds__(256) k_row_rescale(float *__restrict__ A, const float *__restrict__ nrm, int s) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= s) return;
    float *a = A + (size_t)row * s;
    float mn = INFINITY;
    for (int j = lane; j < s; j += 32) mn = fminf(mn, a[j]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    float mx = -INFINITY;
    for (int j = lane; j < s; j += 32) mx = fmaxf(mx, a[j] - mn);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float den = mx + 10e-10f;
    for (int j = lane; j < s; j += 32) a[j] = nrm[j] == 0.f ? 0.f : (a[j] - mn) / den;
}

// Per candidate row of soft masks [m, s]: number of segments >= thr, sum of their soft values (fp64, rounded once), number of
// points (weights w), bounding box over the member segments' boxes.  One warp per row.
__global__ void __launch_bounds__(256)
k_row_stats(const float *__restrict__ soft, int ld, int m, int s, float thr, const int *__restrict__ w, const double *__restrict__ seg_min,
            const double *__restrict__ seg_max, int *__restrict__ count, float *__restrict__ soft_sum, long long *__restrict__ points,
            double *__restrict__ bbox) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= m) return;
    const float *a = soft + (size_t)row * ld;
    int cnt = 0;
    long long pts = 0;
    double sum = 0;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int j = lane; j < s; j += 32) {
        float v = a[j];
        if (v >= thr) {
            ++cnt;
            sum += v;
            if (w) pts += w[j];
            if (seg_min)
#pragma unroll
                for (int c = 0; c < 3; ++c) lo[c] = fmin(lo[c], seg_min[j * 3 + c]), hi[c] = fmax(hi[c], seg_max[j * 3 + c]);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        pts += __shfl_xor_sync(0xffffffffu, pts, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    if (lane == 0) {
        count[row] = cnt;
        soft_sum[row] = (float)sum;
        if (points) points[row] = pts;
        if (bbox)
#pragma unroll
            for (int c = 0; c < 3; ++c) bbox[row * 6 + c] = lo[c], bbox[row * 6 + 3 + c] = hi[c];
    }
}

// inter[i, j] = sum_s w[s] [soft[i,s] >= thr] [soft[j,s] >= thr] — the point-level (mask_i * mask_j).sum() of matrix_nms
__global__ void __launch_bounds__(256)
k_weighted_inter(const float *__restrict__ soft, int ld, int m, int s, float thr, const int *__restrict__ w, int *__restrict__ inter) {
    __shared__ int ta[32][33], tb[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    int acc[4] = {0, 0, 0, 0};
    for (int c0 = 0; c0 < s; c0 += 32) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int rr = ty + 8 * r, c = c0 + tx;
            ta[rr][tx] = (i0 + rr < m && c < s && soft[(size_t)(i0 + rr) * ld + c] >= thr) ? w[c] : 0;
            tb[rr][tx] = (j0 + rr < m && c < s && soft[(size_t)(j0 + rr) * ld + c] >= thr) ? 1 : 0;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            int b = tb[tx][c];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] += ta[ty + 8 * r][c] * b;
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int i = i0 + ty + 8 * r, j = j0 + tx;
        if (i < m && j < m) inter[(size_t)i * m + j] = acc[r];
    }
}

}  // namespace fm
}  // namespace us3d

using namespace us3d;

extern "C" {

int us3d_freemask_soft_masks(const float *f, int s, int d, float *norm, float *soft, void *stream_) {
    US3D_CHECK_ARG(s > 0 && d > 0, "freemask_soft_masks: bad shape");
    cudaStream_t st = (cudaStream_t)stream_;
    fm::k_row_norm<<<ceil_div(s, 8), 256, 0, st>>>(f, s, d, norm);
    US3D_LAUNCH_CHECK();
    fm::k_cosine<<<dim3(ceil_div(s, 32), ceil_div(s, 32)), 256, 0, st>>>(f, norm, s, d, soft);
    US3D_LAUNCH_CHECK();
    fm::k_row_rescale<<<ceil_div(s, 8), 256, 0, st>>>(soft, norm, s);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_freemask_row_stats(const float *soft, int ld, int m, int s, float thr, const int32_t *weights, const double *seg_min,
                            const double *seg_max, int32_t *count, float *soft_sum, long long *points, double *bbox, void *stream_) {
    US3D_CHECK_ARG(m >= 0 && s > 0 && ld >= s, "freemask_row_stats: bad shape");
    US3D_CHECK_ARG((seg_min == nullptr) == (seg_max == nullptr), "freemask_row_stats: seg_min and seg_max go together");
    if (m == 0) return 0;
    fm::k_row_stats<<<ceil_div(m, 8), 256, 0, (cudaStream_t)stream_>>>(soft, ld, m, s, thr, weights, seg_min, seg_max, count, soft_sum,
                                                                         points, bbox);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_freemask_weighted_inter(const float *soft, int ld, int m, int s, float thr, const int32_t *weights, int32_t *inter,
                                 void *stream_) {
    US3D_CHECK_ARG(m >= 0 && s > 0 && ld >= s && weights != nullptr, "freemask_weighted_inter: bad arguments");
    if (m == 0) return 0;
    fm::k_weighted_inter<<<dim3(ceil_div(m, 32), ceil_div(m, 32)), 256, 0, (cudaStream_t)stream_>>>(soft, ld, m, s, thr, weights, inter);
    US3D_LAUNCH_CHECK();
    return 0;
}

// Host-side set logic of freemask_main.py:289-326 (the reference runs it in Python on the host as well): for every candidate
// row of `masks_h` [m, s] (uint8), walk its segments in ascending position and grow / merge blobs over the directed adjacency
// (CSR over positions: adj_ptr_h [s+1], adj_h).  Reproduces the reference's list handling literally, including the skipped slot
// after a merge (`pop` followed by `fused_id += 1`).  Output: blob b belongs to query blob_query_h[b] and consists of
// blob_members_h[blob_ptr_h[b] .. blob_ptr_h[b+1]).  Capacities: max_blobs entries / max_members entries (m*s always suffices).
// Returns the number of blobs, or a negative error.
int us3d_freemask_separate_h(const uint8_t *masks_h, int m, int s, const int32_t *adj_ptr_h, const int32_t *adj_h, int32_t *blob_query_h,
                             int32_t *blob_ptr_h, int32_t *blob_members_h, int max_blobs, long long max_members) {
    US3D_CHECK_ARG(m >= 0 && s > 0, "freemask_separate: bad shape");
    const int words = (s + 63) / 64;
    int nb = 0;
    long long nm = 0;
    blob_ptr_h[0] = 0;
    std::vector<std::vector<uint64_t>> blobs;
    for (int q = 0; q < m; ++q) {
        blobs.clear();
        const uint8_t *row = masks_h + (size_t)q * s;
        for (int c = 0; c < s; ++c) {
            if (!row[c]) continue;
            int last = -1;
            bool merged = false;
            size_t id = 0;
            while (id < blobs.size()) {
                std::vector<uint64_t> &b = blobs[id];
                bool touches = false;
                for (int e = adj_ptr_h[c]; e < adj_ptr_h[c + 1] && !touches; ++e) touches = b[adj_h[e] >> 6] >> (adj_h[e] & 63) & 1ull;
                if (touches) {
                    merged = true;
                    b[c >> 6] |= 1ull << (c & 63);
                    if (last != -1) {
                        for (int w = 0; w < words; ++w) blobs[last][w] |= b[w];
                        blobs.erase(blobs.begin() + id);
                    } else {
                        last = (int)id;
                    }
                }
                ++id;
            }
            if (!merged) {
                blobs.emplace_back(words, 0ull);
                blobs.back()[c >> 6] |= 1ull << (c & 63);
            }
        }
        for (const auto &b : blobs) {
            US3D_CHECK_ARG(nb < max_blobs, "freemask_separate: more than %d blobs", max_blobs);
            blob_query_h[nb] = q;
            for (int c = 0; c < s; ++c)
                if (b[c >> 6] >> (c & 63) & 1ull) {
                    US3D_CHECK_ARG(nm < max_members, "freemask_separate: member capacity exceeded");
                    blob_members_h[nm++] = c;
                }
            blob_ptr_h[++nb] = (int32_t)nm;
        }
    }
    return nb;
}

}  // extern "C"
