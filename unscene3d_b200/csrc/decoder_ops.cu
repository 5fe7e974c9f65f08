// Mask-decoder / matcher helpers: furthest point sampling, segment mean,
// Hungarian cost matrix.
#include "common.cuh"

namespace us3d {

// ---------------------------------------------------------------------------------------------------
// Furthest point sampling.  Restates third_party/pointnet2/_ext_src/src/sampling_gpu.cu:72-176 of the
// reference so that the selected indices are IDENTICAL (mask3d.py:228 feeds integer voxel
// coordinates, so distance ties are the norm):
//   * sample 0 is row 0; rows with x^2+y^2+z^2 <= 1e-3 never update / never win;
//   * thread slot t (T = min(2^floor(log2 n), 512) slots) scans rows t, t+T, ... keeping the first strict
//     maximum of min(d, temp[row]);
//   * the reference's shared-memory tree pairs slot s with s+h for h = T/2 .. 1 and keeps the LOWER
//     position on ties.  Two tied slots first meet at h = lowest set bit of (a xor b) and the one whose
//     bit h is clear survives, i.e. the winner is the tied slot with the smallest BIT-REVERSED index.
//     That is a total order, so any reduction shape that uses the comparator below gives the same row.
struct Cand {
    float d;
    int i;
    unsigned rslot;  // __brev(thread slot)
};

__device__ __forceinline__ Cand fps_better(Cand a, Cand b) {
    if (a.d != b.d) return a.d > b.d ? a : b;
    return a.rslot <= b.rslot ? a : b;
}

__global__ void __launch_bounds__(512) k_fps(const float *__restrict__ xyz, int n, int m, int T, float *__restrict__ temp,
                                             int32_t *__restrict__ idx) {
    if (m <= 0) return;
    __shared__ float sd[16];
    __shared__ int si[16];
    __shared__ unsigned sr[16];
    __shared__ int s_old;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
    xyz += (size_t)blockIdx.x * n * 3;
    temp += (size_t)blockIdx.x * n;
    idx += (size_t)blockIdx.x * m;
    int old = 0;
    if (tid == 0) idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        Cand c{tid < T ? -1.f : -3.f, 0, __brev((unsigned)tid)};
        const float x1 = xyz[old * 3 + 0], y1 = xyz[old * 3 + 1], z1 = xyz[old * 3 + 2];
        if (tid < T)
            for (int k = tid; k < n; k += T) {
                float x2 = xyz[k * 3 + 0], y2 = xyz[k * 3 + 1], z2 = xyz[k * 3 + 2];
                float mag = (x2 * x2) + (y2 * y2) + (z2 * z2);
                if ((double)mag <= 1e-3) continue;
                float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
                float d2 = fminf(d, temp[k]);
                temp[k] = d2;
                if (d2 > c.d) {
                    c.d = d2;
                    c.i = k;
                }
            }
#pragma unroll
        for (int h = 16; h >= 1; h >>= 1) {
            Cand o{__shfl_xor_sync(0xffffffffu, c.d, h), __shfl_xor_sync(0xffffffffu, c.i, h),
                   __shfl_xor_sync(0xffffffffu, c.rslot, h)};
            c = fps_better(c, o);
        }
        if (lane == 0) {
            sd[warp] = c.d;
            si[warp] = c.i;
            sr[warp] = c.rslot;
        }
        __syncthreads();
        if (tid == 0) {
            Cand w{sd[0], si[0], sr[0]};
            for (int q = 1; q < nwarps; ++q) w = fps_better(w, Cand{sd[q], si[q], sr[q]});
            s_old = w.i;
            idx[j] = w.i;
        }
        __syncthreads();
        old = s_old;
    }
}

// ---------------------------------------------------------------------------------------------------
// torch_scatter.scatter_mean over rows
__global__ void __launch_bounds__(256) k_segment_sum(const float *__restrict__ src, const int64_t *__restrict__ index, int n, int c,
                                                     float *out, float *count) {
    long long total = (long long)n * c;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / c), ch = (int)(e % c);
        int s = (int)index[r];
        atomicAdd(&out[(size_t)s * c + ch], src[e]);
        if (ch == 0) atomicAdd(&count[s], 1.f);
    }
}

__global__ void __launch_bounds__(256) k_segment_div(float *out, const float *__restrict__ count, int s, int c) {
    long long total = (long long)s * c;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        float cnt = count[e / c];
        if (cnt > 1.f) out[e] /= cnt;
    }
}

// Exactly rounded variant for the pseudo-mask path: the thresholded affinity (> tau) is decided on these means, so the
// sum is taken in fp64 (order-independent to 2^-53) and rounded to fp32 once — the result does not depend on the
// order in which the atomics land.
__global__ void __launch_bounds__(256) k_segment_sum_f64(const float *__restrict__ src, const int64_t *__restrict__ index, int n, int c,
                                                         double *acc, float *count) {
    long long total = (long long)n * c;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / c), ch = (int)(e % c);
        int s = (int)index[r];
        atomicAdd(&acc[(size_t)s * c + ch], (double)src[e]);
        if (ch == 0) atomicAdd(&count[s], 1.f);
    }
}

__global__ void __launch_bounds__(256) k_segment_div_f64(const double *__restrict__ acc, const float *__restrict__ count, int s, int c,
                                                         float *__restrict__ out) {
    long long total = (long long)s * c;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        float cnt = count[e / c];
        out[e] = (float)(cnt > 1.f ? acc[e] / (double)cnt : acc[e]);
    }
}

__global__ void __launch_bounds__(256) k_segment_mean_bwd(const float *__restrict__ dout, const int64_t *__restrict__ index,
                                                          const float *__restrict__ count, int n, int c, float *__restrict__ dsrc) {
    long long total = (long long)n * c;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / c), ch = (int)(e % c);
        int s = (int)index[r];
        dsrc[e] = dout[(size_t)s * c + ch] / fmaxf(count[s], 1.f);
    }
}

// ---------------------------------------------------------------------------------------------------
// Hungarian cost matrix, one fused pass over logits[s, q] for every target t.
//   block = 32 queries x 8 row lanes, grid = (ceil(q/32), t)
__global__ void __launch_bounds__(256)
k_matcher_cost(const float *__restrict__ logits, int S, int Q, const float *__restrict__ tgt, int T, const float *__restrict__ prob,
               int ncls, const int64_t *__restrict__ labels, float w_class, float w_mask, float w_dice, float *__restrict__ cost) {
    __shared__ float red[5][8][33];
    const int q = blockIdx.x * 32 + threadIdx.x, t = blockIdx.y;
    float bce = 0.f, inter = 0.f, sig_sum = 0.f, tgt_sum = 0.f;
    if (q < Q)
        for (int s = threadIdx.y; s < S; s += 8) {
            float x = logits[(size_t)s * Q + q];
            float z = tgt[(size_t)t * S + s];
            float sp = log1pf(expf(-fabsf(x)));     // softplus(-|x|)
            float pos = fmaxf(-x, 0.f) + sp;        // BCE(x, 1)
            float neg = fmaxf(x, 0.f) + sp;         // BCE(x, 0)
            bce += pos * z + neg * (1.f - z);
            float sg = 1.f / (1.f + expf(-x));
            inter += sg * z;
            sig_sum += sg;
            tgt_sum += z;
        }
    red[0][threadIdx.y][threadIdx.x] = bce;
    red[1][threadIdx.y][threadIdx.x] = inter;
    red[2][threadIdx.y][threadIdx.x] = sig_sum;
    red[3][threadIdx.y][threadIdx.x] = tgt_sum;
    __syncthreads();
    if (threadIdx.y == 0 && q < Q) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 4; ++j) a[j] += red[j][i][threadIdx.x];
        float c_mask = a[0] / (float)S;
        float c_dice = 1.f - (2.f * a[1] + 1.f) / (a[2] + a[3] + 1.f);
        int64_t lab = labels[t];
        float c_class = lab == 253 ? -1.f : -prob[(size_t)q * ncls + (int)lab];
        cost[(size_t)q * T + t] = w_mask * c_mask + w_class * c_class + w_dice * c_dice;
    }
}

static inline int flat_grid2(long long work) {
    long long b = (work + 255) / 256;
    long long cap = (long long)num_sms() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace us3d

using namespace us3d;

extern "C" {

int us3d_furthest_point_sampling(const float *xyz, int b, int n, int m, float *temp, int32_t *idx, void *stream_) {
    US3D_CHECK_ARG(b >= 0 && n > 0 && m >= 0, "fps: bad shape");
    if (b == 0 || m == 0) return 0;
    int slots = 1;
    while (slots * 2 <= n && slots < 512) slots *= 2;  // opt_n_threads of the reference (cuda_utils.h:15-21)
    k_fps<<<b, slots < 32 ? 32 : slots, 0, (cudaStream_t)stream_>>>(xyz, n, m, slots, temp, idx);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_segment_mean_fwd(const float *src, const int64_t *index, int n, int c, int s, float *out, float *count,
                          void *stream_) {
    US3D_CHECK_ARG(n >= 0 && c > 0 && s >= 0, "segment_mean: bad shape");
    if (n == 0 || s == 0) return 0;
    k_segment_sum<<<flat_grid2((long long)n * c), 256, 0, (cudaStream_t)stream_>>>(src, index, n, c, out, count);
    US3D_LAUNCH_CHECK();
    k_segment_div<<<flat_grid2((long long)s * c), 256, 0, (cudaStream_t)stream_>>>(out, count, s, c);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_segment_mean_f64(const float *src, const int64_t *index, int n, int c, int s, double *acc, float *out, float *count,
                          void *stream_) {
    US3D_CHECK_ARG(n >= 0 && c > 0 && s >= 0, "segment_mean_f64: bad shape");
    if (s == 0) return 0;
    if (n > 0) {
        k_segment_sum_f64<<<flat_grid2((long long)n * c), 256, 0, (cudaStream_t)stream_>>>(src, index, n, c, acc, count);
        US3D_LAUNCH_CHECK();
    }
    k_segment_div_f64<<<flat_grid2((long long)s * c), 256, 0, (cudaStream_t)stream_>>>(acc, count, s, c, out);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_segment_mean_bwd(const float *dout, const int64_t *index, const float *count, int n, int c, float *dsrc,
                          void *stream_) {
    US3D_CHECK_ARG(n >= 0 && c > 0, "segment_mean_bwd: bad shape");
    if (n == 0) return 0;
    k_segment_mean_bwd<<<flat_grid2((long long)n * c), 256, 0, (cudaStream_t)stream_>>>(dout, index, count, n, c, dsrc);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_matcher_cost(const float *logits, int s, int q, const float *tgt, int t, const float *prob, int ncls,
                      const int64_t *labels, float w_class, float w_mask, float w_dice, float *cost, void *stream_) {
    US3D_CHECK_ARG(s > 0 && q > 0 && t >= 0 && ncls > 0, "matcher_cost: bad shape");
    if (t == 0) return 0;
    dim3 grid(ceil_div(q, 32), t), block(32, 8);
    k_matcher_cost<<<grid, block, 0, (cudaStream_t)stream_>>>(logits, s, q, tgt, t, prob, ncls, labels, w_class, w_mask, w_dice, cost);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
