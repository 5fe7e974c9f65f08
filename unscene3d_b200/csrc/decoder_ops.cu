// Mask-decoder / matcher helpers: furthest point sampling, segment mean,
// Hungarian cost matrix.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

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
//
// B200 formulation: one thread-block CLUSTER per scene (up to 16 CTAs x 512 threads).  Every row lives on chip for
// the whole sampling loop — 16 rows per thread in registers (x, y, z, running distance), further rows as float4 in
// shared memory, and only what exceeds ~333k rows per scene is streamed from L2 — so a sampling round is a few
// FMAs per row plus one cluster barrier instead of a pass of one CTA over the whole scene.  The winner of a
// round is the maximum of a 64-bit key
//     [ bits(min-distance) + 1 | ~( bitrev(slot) : row / T ) ]          (0 = "slot saw no valid row", as best = -1)
// which orders candidates exactly like the reference's scan + tree; each CTA posts its best key together with
// the row's coordinates into every peer's shared memory (DSMEM), so the next round starts without a global load.
constexpr int kFpsThreads = 512;
constexpr int kFpsRegRows = 16;          // rows per thread kept in registers
constexpr int kFpsMaxSmemSlots = 24;     // rows per thread kept in shared memory (24 * 512 * 16 B = 192 KB)

struct __align__(16) FpsRec {
    unsigned long long key;
    float x, y, z;
    int pad;
};

__device__ __forceinline__ unsigned fps_tie(int k, int lt) {
    // smaller = preferred on equal distance: bit-reversed slot (k mod T) first, then the earlier row of the slot
    unsigned slot = (unsigned)k & ((1u << lt) - 1u);
    unsigned r = lt ? (__brev(slot) >> (32 - lt)) : 0u;
    return (r << 22) | ((unsigned)k >> lt);
}
__device__ __forceinline__ int fps_row_of(unsigned tie, int lt) {
    unsigned r = tie >> 22, q = tie & ((1u << 22) - 1u);
    unsigned slot = lt ? (__brev(r) >> (32 - lt)) : 0u;
    return (int)((q << lt) | slot);
}

// max over the warp of a 64-bit key; returns the lowest lane holding it
__device__ __forceinline__ int fps_warp_argmax(unsigned long long key, unsigned long long &best) {
    unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    unsigned who = __ballot_sync(0xffffffffu, hi == mh && lo == ml);
    best = ((unsigned long long)mh << 32) | ml;
    return __ffs(who) - 1;
}

__global__ void __launch_bounds__(kFpsThreads, 1)
k_fps_cluster(const float *__restrict__ xyz, int n, int m, int lt, int smem_slots, float *__restrict__ temp,
              int32_t *__restrict__ idx) {
    extern __shared__ float4 s_rows[];  // [smem_slots][kFpsThreads]
    __shared__ FpsRec s_warp[kFpsThreads / 32];
    __shared__ FpsRec s_cta[2][16];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stride = C * kFpsThreads, g = rank * kFpsThreads + tid;
    xyz += (size_t)blockIdx.y * n * 3;
    temp += (size_t)blockIdx.y * n;
    idx += (size_t)blockIdx.y * m;

    // ---- load the rows this thread owns: row(i) = i * stride + g
    float rx[kFpsRegRows], ry[kFpsRegRows], rz[kFpsRegRows], rt[kFpsRegRows];
    unsigned valid = 0;
#pragma unroll
    for (int i = 0; i < kFpsRegRows; ++i) {
        int k = i * stride + g;
        rx[i] = ry[i] = rz[i] = 0.f;
        rt[i] = 0.f;
        if (k < n) {
            rx[i] = xyz[(size_t)k * 3 + 0], ry[i] = xyz[(size_t)k * 3 + 1], rz[i] = xyz[(size_t)k * 3 + 2];
            rt[i] = temp[k];
            float mag = (rx[i] * rx[i]) + (ry[i] * ry[i]) + (rz[i] * rz[i]);
            if (!((double)mag <= 1e-3)) valid |= 1u << i;
        }
    }
    for (int j = 0; j < smem_slots; ++j) {
        int k = (kFpsRegRows + j) * stride + g;
        float4 v = make_float4(0.f, 0.f, 0.f, -1.f);  // w < 0: no row / skipped row (distances are >= 0)
        if (k < n) {
            v.x = xyz[(size_t)k * 3 + 0], v.y = xyz[(size_t)k * 3 + 1], v.z = xyz[(size_t)k * 3 + 2];
            float mag = (v.x * v.x) + (v.y * v.y) + (v.z * v.z);
            if (!((double)mag <= 1e-3)) v.w = temp[k];
        }
        s_rows[j * kFpsThreads + tid] = v;
    }
    const int k_stream0 = (kFpsRegRows + smem_slots) * stride + g;  // rows beyond the on-chip capacity

    float x1 = xyz[0], y1 = xyz[1], z1 = xyz[2];
    if (g == 0) idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        unsigned long long best = 0ull;
        float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
        for (int i = 0; i < kFpsRegRows; ++i) {
            if (valid >> i & 1u) {
                float d = (rx[i] - x1) * (rx[i] - x1) + (ry[i] - y1) * (ry[i] - y1) + (rz[i] - z1) * (rz[i] - z1);
                float d2 = fminf(d, rt[i]);
                rt[i] = d2;
                unsigned long long key =
                    ((unsigned long long)(__float_as_uint(d2) + 1u) << 32) | (0xffffffffu - fps_tie(i * stride + g, lt));
                if (key > best) best = key, bx = rx[i], by = ry[i], bz = rz[i];
            }
        }
        for (int s = 0; s < smem_slots; ++s) {
            float4 v = s_rows[s * kFpsThreads + tid];
            if (v.w >= 0.f) {
                float d = (v.x - x1) * (v.x - x1) + (v.y - y1) * (v.y - y1) + (v.z - z1) * (v.z - z1);
                float d2 = fminf(d, v.w);
                s_rows[s * kFpsThreads + tid].w = d2;
                unsigned long long key = ((unsigned long long)(__float_as_uint(d2) + 1u) << 32) |
                                         (0xffffffffu - fps_tie((kFpsRegRows + s) * stride + g, lt));
                if (key > best) best = key, bx = v.x, by = v.y, bz = v.z;
            }
        }
        for (int k = k_stream0; k < n; k += stride) {
            float x2 = xyz[(size_t)k * 3 + 0], y2 = xyz[(size_t)k * 3 + 1], z2 = xyz[(size_t)k * 3 + 2];
            float mag = (x2 * x2) + (y2 * y2) + (z2 * z2);
            if ((double)mag <= 1e-3) continue;
            float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
            float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            unsigned long long key = ((unsigned long long)(__float_as_uint(d2) + 1u) << 32) | (0xffffffffu - fps_tie(k, lt));
            if (key > best) best = key, bx = x2, by = y2, bz = z2;
        }
        // ---- warp -> CTA -> cluster
        unsigned long long wbest;
        int wl = fps_warp_argmax(best, wbest);
        if (lane == wl) s_warp[warp] = FpsRec{best, bx, by, bz, 0};
        __syncthreads();
        if (warp == 0) {
            unsigned long long cbest;
            int ww = fps_warp_argmax(s_warp[lane & (kFpsThreads / 32 - 1)].key, cbest);
            FpsRec rec = s_warp[ww];
            if (lane < C) *cluster.map_shared_rank(&s_cta[j & 1][rank], lane) = rec;
        }
        cluster.sync();
        unsigned long long gbest;
        int wc = fps_warp_argmax(s_cta[j & 1][lane & (C - 1)].key, gbest);
        FpsRec win = s_cta[j & 1][wc];
        int old = 0;
        if ((unsigned)(gbest >> 32) == 0u) {  // no valid row anywhere: the reference's tree returns index 0
            x1 = xyz[0], y1 = xyz[1], z1 = xyz[2];
        } else {
            old = fps_row_of(0xffffffffu - (unsigned)gbest, lt);
            x1 = win.x, y1 = win.y, z1 = win.z;
        }
        if (g == 0) idx[j] = old;
    }
    // ---- the running distances are an in/out argument of the reference op
#pragma unroll
    for (int i = 0; i < kFpsRegRows; ++i) {
        int k = i * stride + g;
        if (k < n) temp[k] = rt[i];
    }
    for (int s = 0; s < smem_slots; ++s) {
        int k = (kFpsRegRows + s) * stride + g;
        float w = s_rows[s * kFpsThreads + tid].w;
        if (k < n && w >= 0.f) temp[k] = w;
    }
    cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
}

// ---------------------------------------------------------------------------------------------------
// Fourier positional encoding (models/position_embedding.py:128-172, get_fourier_embeddings):
//   t = (xyz - lo) / (hi - lo)   [normalize]     t *= 2 pi     proj = t . B[:, :d]     out = [sin(proj) | cos(proj)]
// written row-major [n, 2 d] — the layout every consumer of the reference's [1, 2 d, n] tensor permutes it into
// (models/mask3d.py:195-196, 238-240).  One thread per (row, frequency); same operation order as the reference (shift, scale,
// divide, 2 pi, three-term dot product), accurate sinf / cosf.
__global__ void __launch_bounds__(256)
k_fourier_posenc(const float *__restrict__ xyz, int n, int ld, const float *__restrict__ lo, const float *__restrict__ hi,
                 const float *__restrict__ gauss_b, int ldb, int d, float *__restrict__ out) {
    const long long total = (long long)n * d;
    const float two_pi = 6.283185307179586f;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / d), f = (int)(e - (long long)r * d);
        float proj = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float t = xyz[(size_t)r * ld + a];
            if (lo != nullptr) t = ((t - lo[a]) * 1.0f) / (hi[a] - lo[a]) + 0.0f;
            t *= two_pi;
            proj = fmaf(t, gauss_b[(size_t)a * ldb + f], proj);
        }
        float sn, cs;
        sincosf(proj, &sn, &cs);
        out[(size_t)r * (2 * d) + f] = sn;
        out[(size_t)r * (2 * d) + d + f] = cs;
    }
}

// ---------------------------------------------------------------------------------------------------
// torch_scatter.scatter_mean over rows
// (rows whose index lies outside [0, n_seg) are skipped: the host side validates an explicit dim_size, the kernel never
// writes out of bounds)
__global__ void __launch_bounds__(256) k_segment_sum(const float *__restrict__ src, const int64_t *__restrict__ index, int n, int c,
                                                     int n_seg, float *out, float *count) {
    long long total = (long long)n * c;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / c), ch = (int)(e % c);
        const int64_t s64 = index[r];
        if (s64 < 0 || s64 >= n_seg) continue;
        int s = (int)s64;
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
                                                         int n_seg, double *acc, float *count) {
    long long total = (long long)n * c;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / c), ch = (int)(e % c);
        const int64_t s64 = index[r];
        if (s64 < 0 || s64 >= n_seg) continue;
        int s = (int)s64;
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
        // a label outside [0, ncls) (other than the 253 "mask without class" the reference special-cases, models/matcher.py:
        // 125-127) is an IndexError in the reference; here it poisons the cost so that the assignment solver rejects the matrix
        float c_class = lab == 253 ? -1.f : ((lab < 0 || lab >= ncls) ? __int_as_float(0x7fc00000) : -prob[(size_t)q * ncls + (int)lab]);
        cost[(size_t)q * T + t] = w_mask * c_mask + w_class * c_class + w_dice * c_dice;
    }
}

// ---------------------------------------------------------------------------------------------------
// Mask losses of the set criterion (models/criterion.py:22-73, 168-216): for the matched (query, target) pairs of one scene
//   loss_mask = sum_t w_t * mean_s BCE(x[s, q_t], y[t_t, s]) / n        loss_dice = sum_t w_t * (1 - (2 sum p y + 1) / (sum p + sum y + 1)) / n
// One block per pair reads the pair's logit column and target row once; per-pair sums are kept for backward.
// The reference runs ~20 element-wise / reduction kernels per scene and decoder output for this (x 13 x B per step).
template <typename TT>
__global__ void __launch_bounds__(256)
k_mask_loss_pairs(const float *__restrict__ logits, int S, int Q, const TT *__restrict__ tgt, const int64_t *__restrict__ qidx,
                  const int64_t *__restrict__ tidx, float *__restrict__ stats) {
    __shared__ double red[4][8];
    const int t = blockIdx.x;
    const int q = (int)qidx[t];
    const TT *y = tgt + (size_t)tidx[t] * S;
    double a[4] = {0, 0, 0, 0};
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const float x = logits[(size_t)s * Q + q];
        const float z = (float)y[s];
        a[0] += fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));  // binary_cross_entropy_with_logits
        const float p = 1.f / (1.f + expf(-x));
        a[1] += p * z;
        a[2] += p;
        a[3] += z;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int o = 16; o; o >>= 1) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
        if ((threadIdx.x & 31) == 0) red[j][threadIdx.x >> 5] = a[j];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
        stats[t * 4 + threadIdx.x] = (float)v;
    }
}

// out[0] = loss_mask, out[1] = loss_dice of the scene (fixed summation order)
__global__ void k_mask_loss_finish(const float *__restrict__ stats, const float *__restrict__ weights, int T, int S, float n, float *out) {
    if (threadIdx.x != 0) return;
    float ce = 0.f, dice = 0.f;
    for (int t = 0; t < T; ++t) {
        const float w = weights ? weights[t] : 1.f;
        ce += w * (stats[t * 4] / (float)S);
        dice += w * (1.f - (2.f * stats[t * 4 + 1] + 1.f) / (stats[t * 4 + 2] + stats[t * 4 + 3] + 1.f));
    }
    out[0] = ce / n;
    out[1] = dice / n;
}

// dlogits[s, q_t] = g_ce * w / (n S) * (p - y) - g_dice * w / n * (2 y D - (2 num + 1)) p (1 - p) / D^2 ; other columns stay zero
template <typename TT>
__global__ void __launch_bounds__(256)
k_mask_loss_bwd(const float *__restrict__ logits, int S, int Q, const TT *__restrict__ tgt, const int64_t *__restrict__ qidx,
                const int64_t *__restrict__ tidx, const float *__restrict__ stats, const float *__restrict__ weights, float n,
                const float *__restrict__ gout, float *__restrict__ dlogits) {
    const int t = blockIdx.x;
    const int q = (int)qidx[t];
    const TT *y = tgt + (size_t)tidx[t] * S;
    const float w = weights ? weights[t] : 1.f;
    const float g_ce = gout[0] * w / (n * (float)S), g_dice = gout[1] * w / n;
    const float num2 = 2.f * stats[t * 4 + 1] + 1.f, D = stats[t * 4 + 2] + stats[t * 4 + 3] + 1.f;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const float x = logits[(size_t)s * Q + q];
        const float z = (float)y[s];
        const float p = 1.f / (1.f + expf(-x));
        dlogits[(size_t)s * Q + q] = g_ce * (p - z) - g_dice * (2.f * z * D - num2) * p * (1.f - p) / (D * D);
    }
}

// ---------------------------------------------------------------------------------------------------
// Attention masks of the decoder rounds (models/mask3d.py:407-446) without the point-level detour.  The reference gathers the
// segment logits to every voxel ([sum N, Q] fp32), wraps them in a SparseTensor, average-pools 1..4 times and thresholds.
// Nested average pooling is linear, so the pooled logit of a coarse voxel v is sum_s A[v, s] * seglogit[s, :] with a sparse
// matrix A (CSR, built once per step from the coordinate manager's parent maps, a few entries per row):
//     bits[v, q] = sigmoid( sum_e val[e] * seg[col[e], q] ) < 0.5
__global__ void __launch_bounds__(128)
k_pooled_mask_bits(const int64_t *__restrict__ rowptr, const int64_t *__restrict__ col, const float *__restrict__ val, int n_rows,
                   const float *__restrict__ seg, int Q, uint8_t *__restrict__ bits) {
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const long long e0 = rowptr[r], e1 = rowptr[r + 1];
        for (int q = threadIdx.x; q < Q; q += blockDim.x) {
            float acc = 0.f;
            for (long long e = e0; e < e1; ++e) acc = fmaf(val[e], seg[(size_t)col[e] * Q + q], acc);
            bits[(size_t)r * Q + q] = (1.f / (1.f + expf(-acc))) < 0.5f ? 1 : 0;
        }
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
    int lt = 0;
    while ((2 << lt) <= n && lt < 9) ++lt;  // T = 2^lt slots = opt_n_threads of the reference (cuda_utils.h:15-21)
    static int max_cluster = 0;
    if (!max_cluster) {
        max_cluster = cudaFuncSetAttribute(k_fps_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? 16 : 8;
        US3D_CUDA(cudaFuncSetAttribute(k_fps_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kFpsMaxSmemSlots * kFpsThreads * (int)sizeof(float4)));
        cudaGetLastError();
    }
    int C = 1;
    while (C < max_cluster && (long long)C * kFpsThreads * kFpsRegRows < n) C *= 2;
    long long in_regs = (long long)C * kFpsThreads * kFpsRegRows;
    int smem_slots = n > in_regs ? (int)((n - in_regs + (long long)C * kFpsThreads - 1) / ((long long)C * kFpsThreads)) : 0;
    if (smem_slots > kFpsMaxSmemSlots) smem_slots = kFpsMaxSmemSlots;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C, b, 1);
    cfg.blockDim = dim3(kFpsThreads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem_slots * kFpsThreads * sizeof(float4);
    cfg.stream = (cudaStream_t)stream_;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    US3D_CUDA(cudaLaunchKernelEx(&cfg, k_fps_cluster, xyz, n, m, lt, smem_slots, temp, idx));
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_segment_mean_fwd(const float *src, const int64_t *index, int n, int c, int s, float *out, float *count,
                          void *stream_) {
    US3D_CHECK_ARG(n >= 0 && c > 0 && s >= 0, "segment_mean: bad shape");
    if (n == 0 || s == 0) return 0;
    k_segment_sum<<<flat_grid2((long long)n * c), 256, 0, (cudaStream_t)stream_>>>(src, index, n, c, s, out, count);
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
        k_segment_sum_f64<<<flat_grid2((long long)n * c), 256, 0, (cudaStream_t)stream_>>>(src, index, n, c, s, acc, count);
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

int us3d_fourier_posenc(const float *xyz, int n, int ld, const float *lo, const float *hi, const float *gauss_b, int ldb,
                        int d_out, float *out, void *stream_) {
    US3D_CHECK_ARG(n >= 0 && ld >= 3 && d_out > 0 && ldb >= d_out, "fourier_posenc: bad shape");
    US3D_CHECK_ARG((lo == nullptr) == (hi == nullptr), "fourier_posenc: lo and hi come together");
    if (n == 0) return 0;
    k_fourier_posenc<<<flat_grid2((long long)n * d_out), 256, 0, (cudaStream_t)stream_>>>(xyz, n, ld, lo, hi, gauss_b, ldb, d_out, out);
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

int us3d_mask_loss_fwd(const float *logits, int s, int q, const void *tgt, int tgt_is_float, const int64_t *qidx, const int64_t *tidx,
                       int t, const float *weights, float n, float *stats, float *out, void *stream_) {
    US3D_CHECK_ARG(s > 0 && q > 0 && t >= 0 && n >= 0.f, "mask_loss_fwd: bad shape");  // n == 0 (scene without targets): NaN, as the reference
    cudaStream_t st = (cudaStream_t)stream_;
    if (t > 0) {
        if (tgt_is_float)
            k_mask_loss_pairs<float><<<t, 256, 0, st>>>(logits, s, q, (const float *)tgt, qidx, tidx, stats);
        else
            k_mask_loss_pairs<uint8_t><<<t, 256, 0, st>>>(logits, s, q, (const uint8_t *)tgt, qidx, tidx, stats);
        US3D_LAUNCH_CHECK();
    }
    k_mask_loss_finish<<<1, 32, 0, st>>>(stats, weights, t, s, n, out);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_mask_loss_bwd(const float *logits, int s, int q, const void *tgt, int tgt_is_float, const int64_t *qidx, const int64_t *tidx,
                       int t, const float *weights, float n, const float *stats, const float *gout, float *dlogits, void *stream_) {
    US3D_CHECK_ARG(s > 0 && q > 0 && t >= 0 && n >= 0.f, "mask_loss_bwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CUDA(cudaMemsetAsync(dlogits, 0, sizeof(float) * (size_t)s * q, st));
    if (t == 0) return 0;
    if (tgt_is_float)
        k_mask_loss_bwd<float><<<t, 256, 0, st>>>(logits, s, q, (const float *)tgt, qidx, tidx, stats, weights, n, gout, dlogits);
    else
        k_mask_loss_bwd<uint8_t><<<t, 256, 0, st>>>(logits, s, q, (const uint8_t *)tgt, qidx, tidx, stats, weights, n, gout, dlogits);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_pooled_mask_bits(const int64_t *rowptr, const int64_t *col, const float *val, int n_rows, const float *seg, int q,
                          uint8_t *bits, void *stream_) {
    US3D_CHECK_ARG(n_rows >= 0 && q > 0, "pooled_mask_bits: bad shape");
    if (n_rows == 0) return 0;
    int grid = n_rows < num_sms() * 16 ? n_rows : num_sms() * 16;
    k_pooled_mask_bits<<<grid, 128, 0, (cudaStream_t)stream_>>>(rowptr, col, val, n_rows, seg, q, bits);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
