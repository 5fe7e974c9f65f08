// Fused streaming passes around the sparse convolutions — one launch where the first
// implementation used three to five, and no separate fp32 -> bf16 split pass:
//
//   us3d_bn_stats_fused      column sums + (last block) mean / invstd / running statistics / num_batches_tracked
//   us3d_bn_apply_planes     y = [relu](bn(x) [+ residual]) written as fp32 rows AND as the bf16 hi / lo planes the
//                            tcgen05 convolution of the next layer gathers from
//   us3d_bn_backward_planes  BatchNorm (+ReLU mask) backward; dx written as fp32 rows and bf16 planes (the input
//                            gradient and the weight gradient of the preceding convolution read planes)
//   us3d_spconv_pack_pair    forward and input-gradient weight images in one call
//   us3d_stem_conv_fwd / us3d_stem_conv_wgrad   the 3-channel stem convolution (conv0p1s1, models/res16unet.py:219-221),
//                            whose K = 3 contraction does not fit a tensor-core tile: warp = rows, lane = output channel
//
// Workspaces handed to the BatchNorm kernels are zero on entry and are left zero on exit (the block that finishes
// last cleans up), so no memset launch sits between two layers.
//
// Replaces ME.MinkowskiBatchNorm / MinkowskiReLU / `out += residual` (/root/reference/models/modules/common.py:20-22,
// models/modules/resnet_block.py:48-64) and the stem MinkowskiConvolution.
#include <cuda_bf16.h>

#include "common.cuh"

namespace us3d {
namespace fused {

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ float bf16_residue(float f) { return f - __bfloat162float(__float2bfloat16_rn(f)); }

__device__ __forceinline__ void store_planes(const float (&f)[8], uint4 *hi, uint4 *lo, size_t e) {
    uint4 h;
    h.x = pack_bf16(f[0], f[1]);
    h.y = pack_bf16(f[2], f[3]);
    h.z = pack_bf16(f[4], f[5]);
    h.w = pack_bf16(f[6], f[7]);
    hi[e] = h;
    if (lo != nullptr) {
        uint4 w;
        w.x = pack_bf16(bf16_residue(f[0]), bf16_residue(f[1]));
        w.y = pack_bf16(bf16_residue(f[2]), bf16_residue(f[3]));
        w.z = pack_bf16(bf16_residue(f[4]), bf16_residue(f[5]));
        w.w = pack_bf16(bf16_residue(f[6]), bf16_residue(f[7]));
        lo[e] = w;
    }
}

constexpr int kRows = 256;  // rows per reduction block on large maps
// Small maps (the coarse levels: 513 .. 12k rows) are latency-bound: with 256 rows per block each thread walks 32 dependent-
// issue rows and the whole launch sits at ~10 us however little data there is (ncu launch list: no k_bn_bwd_reduce launch under
// 8 us).  Fewer rows per block = more, shorter blocks; large maps keep 256 to bound the number of double atomics per column.
static inline int reduce_rows(int n) { return n >= 65536 ? kRows : (n >= 16384 ? 128 : (n >= 4096 ? 64 : 32)); }

// ws layout: [0, c) sum, [c, 2c) sum of squares (doubles), then one unsigned ticket counter
struct StatsArgs {
    const float *x;
    int ldx, n, c;
    float eps, momentum;
    float *mean, *invstd, *running_mean, *running_var;
    long long *num_batches_tracked;
    double *ws;
    int rows;  // rows per block
};

__global__ void __launch_bounds__(256) k_bn_stats_fused(StatsArgs a) {
    __shared__ double s0[8][33], s1[8][33];
    __shared__ bool last;
    pdl_wait();
    pdl_trigger();
    const int ch = blockIdx.y * 32 + threadIdx.x;
    const int r_begin = blockIdx.x * a.rows, r_end = min(a.n, r_begin + a.rows);
    float a0 = 0.f, a1 = 0.f;
    if (ch < a.c)
        for (int r = r_begin + threadIdx.y; r < r_end; r += 8) {
            const float v = a.x[(size_t)r * a.ldx + ch];
            a0 += v;
            a1 += v * v;
        }
    s0[threadIdx.y][threadIdx.x] = (double)a0;
    s1[threadIdx.y][threadIdx.x] = (double)a1;
    __syncthreads();
    if (threadIdx.y == 0 && ch < a.c) {
        double t0 = 0, t1 = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            t0 += s0[i][threadIdx.x];
            t1 += s1[i][threadIdx.x];
        }
        atomicAdd(&a.ws[ch], t0);
        atomicAdd(&a.ws[a.c + ch], t1);
    }
    // the block that takes the last ticket sees every other block's sums (fence + atomic), finalises and cleans up
    __threadfence();
    __syncthreads();
    unsigned int *ticket = reinterpret_cast<unsigned int *>(a.ws + 2 * a.c);
    if (threadIdx.x == 0 && threadIdx.y == 0) last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int k = tid; k < a.c; k += 256) {
        const double sum = __ldcg(&a.ws[k]), sq = __ldcg(&a.ws[a.c + k]);
        const double m = sum / a.n;
        double var = sq / a.n - m * m;
        if (var < 0) var = 0;
        a.mean[k] = (float)m;
        a.invstd[k] = (float)(1.0 / sqrt(var + (double)a.eps));
        if (a.running_mean) a.running_mean[k] = (float)((1.0 - a.momentum) * a.running_mean[k] + a.momentum * m);
        if (a.running_var) {
            const double unbiased = a.n > 1 ? var * a.n / (a.n - 1.0) : var;
            a.running_var[k] = (float)((1.0 - a.momentum) * a.running_var[k] + a.momentum * unbiased);
        }
        a.ws[k] = 0.0;
        a.ws[a.c + k] = 0.0;
    }
    if (tid == 0) {
        *ticket = 0u;
        if (a.num_batches_tracked) *a.num_batches_tracked += 1;
    }
}

// one thread = 8 consecutive channels of one row
__global__ void __launch_bounds__(256)
k_bn_apply_planes(const float *__restrict__ x, int ldx, int n, int c, const float *__restrict__ mean, const float *__restrict__ invstd,
                  const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ res, int ldr, int relu,
                  float *__restrict__ y, int ldy, uint4 *__restrict__ hi, uint4 *__restrict__ lo) {
    const int g = c / 8;
    const long long total = (long long)n * g;
    pdl_wait();
    pdl_trigger();
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / g), ch = (int)(e % g) * 8;
        const float4 *src = reinterpret_cast<const float4 *>(x + (size_t)r * ldx + ch);
        const float4 xa = src[0], xb = src[1];
        float f[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        float rs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (res) {
            const float4 *rp = reinterpret_cast<const float4 *>(res + (size_t)r * ldr + ch);
            const float4 ra = rp[0], rb = rp[1];
            rs[0] = ra.x; rs[1] = ra.y; rs[2] = ra.z; rs[3] = ra.w; rs[4] = rb.x; rs[5] = rb.y; rs[6] = rb.z; rs[7] = rb.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float o = (f[i] - __ldg(mean + ch + i)) * __ldg(invstd + ch + i) * __ldg(gamma + ch + i) + __ldg(beta + ch + i);
            if (res) o += rs[i];
            if (relu) o = fmaxf(o, 0.f);
            f[i] = o;
        }
        float4 *dst = reinterpret_cast<float4 *>(y + (size_t)r * ldy + ch);
        dst[0] = make_float4(f[0], f[1], f[2], f[3]);
        dst[1] = make_float4(f[4], f[5], f[6], f[7]);
        if (hi != nullptr) store_planes(f, hi, lo, (size_t)e);
    }
}

// column sums of g and g * xhat (g = dy masked by the ReLU), block = 32 channels x 8 row lanes
__global__ void __launch_bounds__(256)
k_bn_bwd_reduce(const float *__restrict__ dy, int lddy, const float *__restrict__ x, int ldx, const float *__restrict__ y, int ldy, int n,
                int c, const float *__restrict__ mean, const float *__restrict__ invstd, int relu, double *red, int rows) {
    __shared__ double s0[8][33], s1[8][33];
    const int ch = blockIdx.y * 32 + threadIdx.x;
    const int r_begin = blockIdx.x * rows, r_end = min(n, r_begin + rows);
    float a0 = 0.f, a1 = 0.f;
    pdl_wait();
    pdl_trigger();
    if (ch < c) {
        const float m = mean[ch], is = invstd[ch];
        for (int r = r_begin + threadIdx.y; r < r_end; r += 8) {
            float g = dy[(size_t)r * lddy + ch];
            if (relu && !(y[(size_t)r * ldy + ch] > 0.f)) g = 0.f;
            a0 += g;
            a1 += g * ((x[(size_t)r * ldx + ch] - m) * is);
        }
    }
    s0[threadIdx.y][threadIdx.x] = (double)a0;
    s1[threadIdx.y][threadIdx.x] = (double)a1;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
        double t0 = 0, t1 = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            t0 += s0[i][threadIdx.x];
            t1 += s1[i][threadIdx.x];
        }
        atomicAdd(&red[ch], t0);
        atomicAdd(&red[c + ch], t1);
    }
}

// W = 8 channels per thread (vector path with planes) or 1 (any shape).  The per-channel coefficients of
//     dx = a[ch] * g + b[ch] * x + d[ch]        (a = invstd gamma, b = -a invstd s1 / n, d = -a s0 / n - b mean)
// are folded once per block into shared memory (the double-precision column sums would otherwise be re-read from L2
// by every thread for every element).
template <int W>
__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const float *__restrict__ dy, int lddy, const float *__restrict__ x, int ldx, const float *__restrict__ y, int ldy, int n,
               int c, const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma, int relu,
               double *red, float *__restrict__ dx, int lddx, float *__restrict__ dres, int lddres, float *dgamma, float *dbeta,
               int batch_terms, uint4 *__restrict__ dx_hi, uint4 *__restrict__ dx_lo) {
    extern __shared__ float coef[];  // [3][c]
    __shared__ bool last;
    float *ca = coef, *cb = coef + c, *cd = coef + 2 * c;
    pdl_wait();
    pdl_trigger();
    {
        const double inv_n = 1.0 / (double)n;
        for (int k = threadIdx.x; k < c; k += blockDim.x) {
            const double is = (double)invstd[k], m = (double)mean[k];
            const double a = is * (double)gamma[k];
            const double s0 = batch_terms ? red[k] : 0.0, s1 = batch_terms ? red[c + k] : 0.0;
            const double b = -a * is * s1 * inv_n;
            ca[k] = (float)a;
            cb[k] = (float)b;
            cd[k] = (float)(-a * s0 * inv_n - b * m);
        }
    }
    __syncthreads();
    const int g8 = c / W;
    const long long total = (long long)n * g8;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / g8), ch = (int)(e % g8) * W;
        float gv[W], xv[W], yv[W], o[W];
        if constexpr (W == 8) {
            const float4 *gp = reinterpret_cast<const float4 *>(dy + (size_t)r * lddy + ch);
            const float4 *xp = reinterpret_cast<const float4 *>(x + (size_t)r * ldx + ch);
            const float4 ga = gp[0], gb = gp[1], xa = xp[0], xb = xp[1];
            gv[0] = ga.x; gv[1] = ga.y; gv[2] = ga.z; gv[3] = ga.w; gv[4] = gb.x; gv[5] = gb.y; gv[6] = gb.z; gv[7] = gb.w;
            xv[0] = xa.x; xv[1] = xa.y; xv[2] = xa.z; xv[3] = xa.w; xv[4] = xb.x; xv[5] = xb.y; xv[6] = xb.z; xv[7] = xb.w;
            if (relu) {
                const float4 *yp = reinterpret_cast<const float4 *>(y + (size_t)r * ldy + ch);
                const float4 ya = yp[0], yb = yp[1];
                yv[0] = ya.x; yv[1] = ya.y; yv[2] = ya.z; yv[3] = ya.w; yv[4] = yb.x; yv[5] = yb.y; yv[6] = yb.z; yv[7] = yb.w;
            }
        } else {
            gv[0] = dy[(size_t)r * lddy + ch];
            xv[0] = x[(size_t)r * ldx + ch];
            if (relu) yv[0] = y[(size_t)r * ldy + ch];
        }
#pragma unroll
        for (int i = 0; i < W; ++i) {
            float g = gv[i];
            if (relu && !(yv[i] > 0.f)) g = 0.f;
            gv[i] = g;
            o[i] = fmaf(ca[ch + i], g, fmaf(cb[ch + i], xv[i], cd[ch + i]));
        }
        if constexpr (W == 8) {
            float4 *dp = reinterpret_cast<float4 *>(dx + (size_t)r * lddx + ch);
            dp[0] = make_float4(o[0], o[1], o[2], o[3]);
            dp[1] = make_float4(o[4], o[5], o[6], o[7]);
            if (dres) {
                float4 *rp = reinterpret_cast<float4 *>(dres + (size_t)r * lddres + ch);
                rp[0] = make_float4(gv[0], gv[1], gv[2], gv[3]);
                rp[1] = make_float4(gv[4], gv[5], gv[6], gv[7]);
            }
            if (dx_hi != nullptr) store_planes(o, dx_hi, dx_lo, (size_t)e);
        } else {
            dx[(size_t)r * lddx + ch] = o[0];
            if (dres) dres[(size_t)r * lddres + ch] = gv[0];
        }
    }
    // the block that finishes last has seen every other block read the sums: publish dgamma / dbeta, zero the workspace
    __threadfence();
    __syncthreads();
    unsigned int *ticket = reinterpret_cast<unsigned int *>(red + 2 * c);
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    for (int k = threadIdx.x; k < c; k += blockDim.x) {
        if (dgamma) dgamma[k] = (float)red[c + k];
        if (dbeta) dbeta[k] = (float)red[k];
        red[k] = 0.0;
        red[c + k] = 0.0;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

// ------------------------------------------------------------------------------------------------- stem convolution
// warp = a strip of output rows, lane = output channel (32 per blockIdx.y slice); the [kvol, CIN, 32] weight slice sits in
// shared memory.  For every row, lane l first fetches the neighbour index of offset l and that neighbour's CIN input
// values (one round of L2 latency for all kvol <= 32 offsets instead of kvol dependent round trips); the contraction
// then broadcasts them with warp shuffles.  Absent neighbours contribute zeros.
template <int CIN>
__global__ void __launch_bounds__(256)
k_stem_fwd(const float *__restrict__ x, int ldx, const int32_t *__restrict__ nbr, int n_rows, int kvol, const float *__restrict__ w, int cout,
           const float *__restrict__ bias, float *__restrict__ y, int ldy) {
    extern __shared__ float ws[];  // [kvol][CIN][32]
    const int co0 = blockIdx.y * 32, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kvol * CIN * 32; i += blockDim.x) {
        const int l = i & 31, kc = i >> 5;
        ws[i] = co0 + l < cout ? w[(size_t)kc * cout + co0 + l] : 0.f;
    }
    __syncthreads();
    const int co = co0 + lane;
    const float b = (bias != nullptr && co < cout) ? bias[co] : 0.f;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r0 = (blockIdx.x * (blockDim.x >> 5) + warp) * 2; r0 < n_rows; r0 += warps * 2) {
        const bool two = r0 + 1 < n_rows;
        float xa[CIN], xb[CIN];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xa[ci] = xb[ci] = 0.f;
        if (lane < kvol) {
            const int i0 = __ldg(nbr + (size_t)lane * n_rows + r0);
            const int i1 = two ? __ldg(nbr + (size_t)lane * n_rows + r0 + 1) : -1;
            if (i0 >= 0) {
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) xa[ci] = __ldg(x + (size_t)i0 * ldx + ci);
            }
            if (i1 >= 0) {
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) xb[ci] = __ldg(x + (size_t)i1 * ldx + ci);
            }
        }
        float acc0 = b, acc1 = b;
        for (int k = 0; k < kvol; ++k) {
            const float *wk = ws + k * CIN * 32 + lane;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                const float wv = wk[ci * 32];
                acc0 = fmaf(__shfl_sync(0xffffffffu, xa[ci], k), wv, acc0);
                acc1 = fmaf(__shfl_sync(0xffffffffu, xb[ci], k), wv, acc1);
            }
        }
        if (co < cout) {
            y[(size_t)r0 * ldy + co] = acc0;
            if (two) y[(size_t)(r0 + 1) * ldy + co] = acc1;
        }
    }
}

// Weight gradient: lane = output channel, each thread keeps the kvol x CIN partial sums of its channel in registers
// over the warp's strip of rows (same lane-parallel fetch + shuffle broadcast as the forward); warps of a block meet
// in shared memory, blocks through one atomic per weight.
template <int CIN, int KVOL>
__global__ void __launch_bounds__(256)
k_stem_wgrad(const float *__restrict__ x, int ldx, const int32_t *__restrict__ nbr, int n_rows, const float *__restrict__ dy, int lddy,
             int cout, float *__restrict__ dw) {
    extern __shared__ float red[];  // [KVOL * CIN][32]
    const int co0 = blockIdx.y * 32, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int co = co0 + lane;
    float acc[KVOL][CIN];
#pragma unroll
    for (int k = 0; k < KVOL; ++k)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) acc[k][ci] = 0.f;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = blockIdx.x * (blockDim.x >> 5) + warp; r < n_rows; r += warps) {
        const float g = co < cout ? __ldg(dy + (size_t)r * lddy + co) : 0.f;
        float xv[CIN];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xv[ci] = 0.f;
        if (lane < KVOL) {
            const int i = __ldg(nbr + (size_t)lane * n_rows + r);
            if (i >= 0) {
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) xv[ci] = __ldg(x + (size_t)i * ldx + ci);
            }
        }
#pragma unroll
        for (int k = 0; k < KVOL; ++k)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) acc[k][ci] = fmaf(__shfl_sync(0xffffffffu, xv[ci], k), g, acc[k][ci]);
    }
    for (int i = threadIdx.x; i < KVOL * CIN * 32; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KVOL; ++k)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) atomicAdd(&red[(k * CIN + ci) * 32 + lane], acc[k][ci]);
    __syncthreads();
    for (int i = threadIdx.x; i < KVOL * CIN * 32; i += blockDim.x) {
        const int l = i & 31, kc = i >> 5;
        if (co0 + l < cout) atomicAdd(&dw[(size_t)kc * cout + co0 + l], red[i]);
    }
}

// ------------------------------------------------------------------------------------------------- weight images
// Both weight images of a convolution in one launch: blockIdx.y = 0 packs the forward image (K = cin, N = cout),
// blockIdx.y = 1 the input-gradient image (K = cout, N = cin, W[k]^T, offsets reversed when `flip`).  Layout as
// us3d_spconv_pack_weights: per (offset, 64-channel chunk, plane) an [N][128 B] slab in the K-major SWIZZLE_128B image.
struct PackSide {
    uint8_t *out;
    int kdim, ndim, transpose, flip;
};

__global__ void __launch_bounds__(256) k_pack_pair(const float *__restrict__ w, int kvol, int w_cin, int w_cout, int planes, PackSide s0,
                                                   PackSide s1) {
    const PackSide s = blockIdx.y == 0 ? s0 : s1;
    if (s.out == nullptr) return;
    const int nchunks = (s.kdim + 63) / 64;
    const long long total = (long long)kvol * nchunks * s.ndim * 8;
    const size_t slab = (size_t)s.ndim * 128;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(e % s.ndim);
        long long t = e / s.ndim;
        const int g = (int)(t % 8);
        t /= 8;
        const int c = (int)(t % nchunks);
        const int k = (int)(t / nchunks);
        const int kq = s.flip ? kvol - 1 - k : k;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kk = c * 64 + g * 8 + i;
            float v = 0.f;
            if (kk < s.kdim) v = s.transpose ? w[((size_t)kq * w_cin + n) * w_cout + kk] : w[((size_t)kq * w_cin + kk) * w_cout + n];
            f[i] = v;
        }
        uint8_t *base = s.out + ((size_t)k * nchunks + c) * planes * slab;
        const size_t off = (size_t)n * 128 + (size_t)((g ^ (n & 7)) << 4);
        store_planes(f, reinterpret_cast<uint4 *>(base + off), planes == 2 ? reinterpret_cast<uint4 *>(base + slab + off) : nullptr, 0);
    }
}

// Every weight image of a network in ONE launch (a training step re-packs all convolutions after the optimizer update:
// 62 launches for Res16UNet34C otherwise).  blockIdx.y = image; descriptors live in device memory.  `w_ci0` / the image's own
// kdim / ndim select a column slice of the input channels (input gradients wider than 256 channels are produced in slices).
struct PackItem {
    const float *w;      // [kvol, w_cin, w_cout] fp32
    uint8_t *out;        // image, layout as us3d_spconv_pack_weights
    int kvol, w_cin, w_cout, w_ci0;
    int kdim, ndim, transpose, flip;
};
static_assert(sizeof(PackItem) == 48, "PackItem is mirrored by a numpy dtype in engine/functional.py");

__global__ void __launch_bounds__(256) k_pack_many(const PackItem *__restrict__ items, int planes) {
    const PackItem s = items[blockIdx.y];
    const int nchunks = (s.kdim + 63) / 64;
    const long long total = (long long)s.kvol * nchunks * s.ndim * 8;
    const size_t slab = (size_t)s.ndim * 128;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(e % s.ndim);
        long long t = e / s.ndim;
        const int g = (int)(t % 8);
        t /= 8;
        const int c = (int)(t % nchunks);
        const int k = (int)(t / nchunks);
        const int kq = s.flip ? s.kvol - 1 - k : k;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kk = c * 64 + g * 8 + i;
            float v = 0.f;
            // forward: K = input channel (kk), N = output channel (n);  gradient: K = output channel, N = input channel
            if (kk < s.kdim)
                v = s.transpose ? s.w[((size_t)kq * s.w_cin + s.w_ci0 + n) * s.w_cout + kk]
                                : s.w[((size_t)kq * s.w_cin + s.w_ci0 + kk) * s.w_cout + n];
            f[i] = v;
        }
        uint8_t *base = s.out + ((size_t)k * nchunks + c) * planes * slab;
        const size_t off = (size_t)n * 128 + (size_t)((g ^ (n & 7)) << 4);
        store_planes(f, reinterpret_cast<uint4 *>(base + off), planes == 2 ? reinterpret_cast<uint4 *>(base + slab + off) : nullptr, 0);
    }
}

// fp32 rows -> bf16 hi plane (+ lo plane = x - hi), row-major [n, c], c % 8 == 0 — only for tensors that come from outside the
// fused passes (network input, loss gradient): activations get their planes from the pass that produces them.
__global__ void __launch_bounds__(256) k_split_bf16(const float *__restrict__ x, int ldx, int n, int c, uint4 *__restrict__ hi,
                                                    uint4 *__restrict__ lo) {
    const int g = c / 8;
    const long long total = (long long)n * g;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / g), q = (int)(e % g);
        const float4 *src = reinterpret_cast<const float4 *>(x + (size_t)r * ldx + q * 8);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        store_planes(f, hi, lo, (size_t)e);
    }
}

static inline int flat_grid(long long work) {
    long long b = (work + 255) / 256;
    long long cap = (long long)num_sms() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace fused
}  // namespace us3d

using namespace us3d;

extern "C" {

int us3d_bn_workspace_bytes(int c) { return (int)(sizeof(double) * 2 * (size_t)c + 16); }

int us3d_bn_stats_fused(const float *x, int ldx, int n, int c, float eps, float momentum, float *mean, float *invstd,
                        float *running_mean, float *running_var, long long *num_batches_tracked, double *ws, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n > 0 && c > 0 && ldx >= c, "bn_stats_fused: bad shape");
    const int rows = fused::reduce_rows(n);
    fused::StatsArgs a{x, ldx, n, c, eps, momentum, mean, invstd, running_mean, running_var, num_batches_tracked, ws, rows};
    dim3 grid(ceil_div(n, rows), ceil_div(c, 32)), block(32, 8);
    US3D_CUDA(launch_pdl(fused::k_bn_stats_fused, grid, block, 0, st, a));
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_bn_apply_planes(const float *x, int ldx, int n, int c, const float *mean, const float *invstd, const float *gamma,
                         const float *beta, const float *residual, int ldr, int relu, float *y, int ldy, void *hi, void *lo,
                         void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n >= 0 && c > 0 && ldx >= c && ldy >= c, "bn_apply_planes: bad shape");
    if (n == 0) return 0;
    const bool vec = c % 8 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && fused::al16(x) && fused::al16(y) &&
                     (!residual || (ldr % 4 == 0 && fused::al16(residual)));
    US3D_CHECK_ARG(vec || hi == nullptr, "bn_apply_planes: planes need c %% 8 == 0 and 16-byte aligned rows");
    if (!vec) return us3d_bn_apply(x, ldx, n, c, mean, invstd, gamma, beta, residual, ldr, relu, y, ldy, stream_);
    US3D_CUDA(launch_pdl(fused::k_bn_apply_planes, dim3(fused::flat_grid((long long)n * c / 8)), dim3(256), 0, st, x, ldx, n, c, mean, invstd, gamma,
                         beta, residual, ldr, relu, y, ldy, (uint4 *)hi, (uint4 *)lo));
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_bn_backward_planes(const float *dy, int lddy, const float *x, int ldx, const float *y, int ldy, int n, int c,
                            const float *mean, const float *invstd, const float *gamma, int relu, int batch_terms, double *ws,
                            float *dx, int lddx, float *dres, int lddres, float *dgamma, float *dbeta, void *dx_hi, void *dx_lo,
                            void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n > 0 && c > 0 && c <= 4096, "bn_backward_planes: bad shape");
    const int rows = fused::reduce_rows(n);
    dim3 grid(ceil_div(n, rows), ceil_div(c, 32)), block(32, 8);
    US3D_CUDA(launch_pdl(fused::k_bn_bwd_reduce, grid, block, 0, st, dy, lddy, x, ldx, y, ldy, n, c, mean, invstd, relu, ws, rows));
    US3D_LAUNCH_CHECK();
    const bool vec = c % 8 == 0 && lddy % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && lddx % 4 == 0 && fused::al16(dy) && fused::al16(x) &&
                     fused::al16(y) && fused::al16(dx) && (!dres || (lddres % 4 == 0 && fused::al16(dres)));
    US3D_CHECK_ARG(vec || dx_hi == nullptr, "bn_backward_planes: planes need c %% 8 == 0 and 16-byte aligned rows");
    if (vec)
        US3D_CUDA(launch_pdl(fused::k_bn_bwd_apply<8>, dim3(fused::flat_grid((long long)n * c / 8)), dim3(256), 3 * (size_t)c * sizeof(float), st,
                             dy, lddy, x, ldx, y, ldy, n, c, mean, invstd, gamma, relu, ws, dx, lddx, dres, lddres, dgamma, dbeta, batch_terms,
                             (uint4 *)dx_hi, (uint4 *)dx_lo));
    else
        fused::k_bn_bwd_apply<1><<<fused::flat_grid((long long)n * c), 256, 3 * (size_t)c * sizeof(float), st>>>(
            dy, lddy, x, ldx, y, ldy, n, c, mean, invstd, gamma, relu, ws, dx, lddx, dres, lddres, dgamma, dbeta, batch_terms, nullptr,
            nullptr);
    US3D_LAUNCH_CHECK();
    return 0;
}

long long us3d_spconv_packed_bytes(int kvol, int kdim, int ndim, int passes) {
    return (long long)kvol * ceil_div(kdim, 64) * (passes == 3 ? 2 : 1) * ndim * 128;
}

int us3d_spconv_tc_supported(int cin, int cout) {
    return cin >= 16 && cin % 16 == 0 && cout >= 16 && cout % 16 == 0 && cout <= 256;
}

int us3d_spconv_wgrad_tc_supported(int cin, int cout) { return cin >= 8 && cin % 8 == 0 && cout >= 16 && cout % 16 == 0 && cout <= 256; }

int us3d_split_bf16(const float *x, int ldx, int n, int c, void *hi, void *lo, void *stream_) {
    US3D_CHECK_ARG(n >= 0 && c > 0 && c % 8 == 0 && ldx % 4 == 0 && ldx >= c, "split_bf16: need c %% 8 == 0 and ldx %% 4 == 0");
    US3D_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "split_bf16: x must be 16-byte aligned");
    if (n == 0) return 0;
    fused::k_split_bf16<<<fused::flat_grid((long long)n * (c / 8)), 256, 0, (cudaStream_t)stream_>>>(x, ldx, n, c, (uint4 *)hi, (uint4 *)lo);
    US3D_LAUNCH_CHECK();
    return 0;
}

/* one image: forward (transpose = 0) or input-gradient (transpose = 1, optionally with the offsets flipped) */
int us3d_spconv_pack_weights(const float *w, int kvol, int cin, int cout, int transpose, int flip_k, int passes, void *out,
                             void *stream_) {
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "pack_weights: kvol %d out of range", kvol);
    US3D_CHECK_ARG(passes == 1 || passes == 3, "pack_weights: passes must be 1 or 3");
    if (transpose)
        return us3d_spconv_pack_pair(w, kvol, cin, cout, flip_k, passes, nullptr, out, stream_);
    US3D_CHECK_ARG(flip_k == 0, "pack_weights: flipped offsets belong to the input-gradient image");
    return us3d_spconv_pack_pair(w, kvol, cin, cout, 0, passes, out, nullptr, stream_);
}

int us3d_spconv_pack_pair(const float *w, int kvol, int cin, int cout, int flip_dgrad, int passes, void *out_fwd, void *out_dgrad,
                          void *stream_) {
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "pack_pair: kvol %d out of range", kvol);
    US3D_CHECK_ARG(passes == 1 || passes == 3, "pack_pair: passes must be 1 or 3");
    US3D_CHECK_ARG(out_fwd == nullptr || us3d_spconv_tc_supported(cin, cout), "pack_pair: unsupported forward shape %d -> %d", cin, cout);
    US3D_CHECK_ARG(out_dgrad == nullptr || us3d_spconv_tc_supported(cout, cin), "pack_pair: unsupported gradient shape %d -> %d", cout, cin);
    if (out_fwd == nullptr && out_dgrad == nullptr) return 0;
    fused::PackSide s0{(uint8_t *)out_fwd, cin, cout, 0, 0}, s1{(uint8_t *)out_dgrad, cout, cin, 1, flip_dgrad};
    const int big = cin > cout ? cin : cout;
    const long long total = (long long)kvol * ceil_div(big, 64) * big * 8;
    int gx = (int)((total + 255) / 256);
    if (gx > num_sms() * 8) gx = num_sms() * 8;
    dim3 grid(gx, 2);
    fused::k_pack_pair<<<grid, 256, 0, (cudaStream_t)stream_>>>(w, kvol, cin, cout, passes == 3 ? 2 : 1, s0, s1);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_spconv_pack_many(const void *items, int n_items, int passes, void *stream_) {
    US3D_CHECK_ARG(n_items >= 0 && n_items <= 65535, "pack_many: %d images", n_items);
    US3D_CHECK_ARG(passes == 1 || passes == 3, "pack_many: passes must be 1 or 3");
    if (n_items == 0) return 0;
    fused::k_pack_many<<<dim3(96, n_items), 256, 0, (cudaStream_t)stream_>>>((const fused::PackItem *)items, passes == 3 ? 2 : 1);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_stem_conv_supported(int cin, int cout, int kvol) { return (cin >= 1 && cin <= 4 && cout >= 1 && kvol >= 1 && kvol <= US3D_MAX_KVOL) ? 1 : 0; }

int us3d_stem_conv_fwd(const float *x, int ldx, const int32_t *nbr, int n_rows, int kvol, const float *w, int cin, int cout,
                       const float *bias, float *y, int ldy, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(us3d_stem_conv_supported(cin, cout, kvol), "stem_conv_fwd: unsupported shape %d -> %d, kvol %d", cin, cout, kvol);
    if (n_rows == 0) return 0;
    int gx = ceil_div(n_rows, 8 * 2 * 4);
    if (gx > num_sms() * 8) gx = num_sms() * 8;
    dim3 grid(gx, ceil_div(cout, 32));
    const size_t smem = (size_t)kvol * cin * 32 * sizeof(float);
    ProfScope prof(st, 0, n_rows, n_rows, kvol, cin, cout);
    switch (cin) {
        case 1: fused::k_stem_fwd<1><<<grid, 256, smem, st>>>(x, ldx, nbr, n_rows, kvol, w, cout, bias, y, ldy); break;
        case 2: fused::k_stem_fwd<2><<<grid, 256, smem, st>>>(x, ldx, nbr, n_rows, kvol, w, cout, bias, y, ldy); break;
        case 3: fused::k_stem_fwd<3><<<grid, 256, smem, st>>>(x, ldx, nbr, n_rows, kvol, w, cout, bias, y, ldy); break;
        default: fused::k_stem_fwd<4><<<grid, 256, smem, st>>>(x, ldx, nbr, n_rows, kvol, w, cout, bias, y, ldy); break;
    }
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_stem_conv_wgrad(const float *x, int ldx, const int32_t *nbr, int n_rows, int kvol, const float *dy, int lddy, float *dw,
                         int cin, int cout, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(us3d_stem_conv_supported(cin, cout, kvol) && (kvol == 27 || kvol == 8 || kvol == 1) && cin == 3,
                   "stem_conv_wgrad: unsupported shape %d -> %d, kvol %d", cin, cout, kvol);
    if (n_rows == 0) return 0;
    int gx = num_sms() * 2;
    if (gx > ceil_div(n_rows, 8)) gx = ceil_div(n_rows, 8);
    dim3 grid(gx, ceil_div(cout, 32));
    const size_t smem = (size_t)kvol * cin * 32 * sizeof(float);
    ProfScope prof(st, 1, n_rows, n_rows, kvol, cin, cout);
    if (kvol == 27)
        fused::k_stem_wgrad<3, 27><<<grid, 256, smem, st>>>(x, ldx, nbr, n_rows, dy, lddy, cout, dw);
    else if (kvol == 8)
        fused::k_stem_wgrad<3, 8><<<grid, 256, smem, st>>>(x, ldx, nbr, n_rows, dy, lddy, cout, dw);
    else
        fused::k_stem_wgrad<3, 1><<<grid, 256, smem, st>>>(x, ldx, nbr, n_rows, dy, lddy, cout, dw);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
