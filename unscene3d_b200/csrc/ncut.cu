// Pseudo-mask NCut step: segment-feature affinity, thresholded {1, eps} graph as a bit
// matrix, and the matrix-vector product that drives the spectral (Fiedler-vector) solve.
//
// Reference: pseudo_masks/unscene3d_pseudo_main.py  normalize_mat :82-86, get_affinity_matrix :89-119,
// second_smallest_eigenvector :138-146 (scipy.linalg.eigh(D - A, D, subset_by_index=[1, 2]) — an O(S^3) dense LAPACK
// solve on the host, x up to 20 iterations per scene).  Here the thresholded affinity never exists as floats: it is
// W = eps * 11^T + (1 - eps) * B with B a bit matrix (S^2 / 8 bytes, L2-resident), so W x is a popcount-free masked
// sum over 32-bit words, and the eigenvector comes from a Lanczos iteration over M = D^-1/2 W D^-1/2 in fp64.
#include "common.cuh"

namespace us3d {

// order-preserving float <-> uint map for atomicMin / atomicMax on floats
__device__ __forceinline__ unsigned f2o(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// row L2 norms (F.normalize: x / max(||x||, 1e-12))
__global__ void __launch_bounds__(256) k_row_inv_norm(const float *__restrict__ f, int s, int d, float *__restrict__ inv) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= s) return;
    double acc = 0;
    for (int c = lane; c < d; c += 32) {
        double v = f[(size_t)row * d + c];
        acc += v * v;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) inv[row] = (float)(1.0 / fmax(sqrt(acc), 1e-12));
}

// A[i, j] = <f_i, f_j> * inv_i * inv_j  (fp32 products, fp32 accumulation like the reference's fp32 GEMM), 32x32 tiles;
// also the global statistics normalize_mat needs: any positive, min over non-zero entries, max.
__global__ void __launch_bounds__(256) k_gram(const float *__restrict__ f, const float *__restrict__ inv, int s, int d,
                                              float *__restrict__ A, unsigned *stats) {
    __shared__ float ta[32][33], tb[32][33];
    __shared__ unsigned s_min, s_max, s_pos;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    if (threadIdx.x == 0) {
        s_min = 0xFFFFFFFFu;
        s_max = 0u;
        s_pos = 0u;
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < d; c0 += 32) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int rr = ty + 8 * r;
            int c = c0 + tx;
            ta[rr][tx] = (i0 + rr < s && c < d) ? f[(size_t)(i0 + rr) * d + c] * inv[i0 + rr] : 0.f;
            tb[rr][tx] = (j0 + rr < s && c < d) ? f[(size_t)(j0 + rr) * d + c] * inv[j0 + rr] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            float b = tb[tx][c];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(ta[ty + 8 * r][c], b, acc[r]);
        }
    }
    unsigned lmin = 0xFFFFFFFFu, lmax = 0u, lpos = 0u;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int i = i0 + ty + 8 * r, j = j0 + tx;
        if (i < s && j < s) {
            float v = acc[r];
            A[(size_t)i * s + j] = v;
            unsigned o = f2o(v);
            if (v != 0.f) lmin = min(lmin, o);
            lmax = max(lmax, o);
            lpos |= v > 0.f;
        }
    }
    __syncthreads();
    atomicMin(&s_min, lmin);
    atomicMax(&s_max, lmax);
    atomicOr(&s_pos, lpos);
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicMin(&stats[0], s_min);
        atomicMax(&stats[1], s_max);
        atomicOr(&stats[2], s_pos);
    }
}

__device__ __forceinline__ float normalize_entry(float a, const unsigned *st) {
    // normalize_mat: A -= min(A[A != 0]) if any(A > 0); A[A < 0] = 0; A /= A.max() + 1e-5   (float32, as numpy does)
    float m = st[2] ? o2f(st[0]) : 0.f;
    float mx = fmaxf(o2f(st[1]) - m, 0.f);
    float v = fmaxf(a - m, 0.f);
    return v / (mx + 1e-5f);
}

// one warp per (row, 32-column word): threshold the (averaged) normalised affinities into a bit word;
// degree from the UNPAINTED graph (the reference computes D before painting, :116-118 vs :426-427),
// bits of painted rows / columns cleared afterwards.
__global__ void __launch_bounds__(256)
k_threshold(const float *__restrict__ Aa, const float *__restrict__ Ab, int s, int words, const unsigned *stats_a,
            const unsigned *stats_b, float tau, double eps, const uint8_t *__restrict__ painted, uint32_t *__restrict__ bits,
            double *degree) {
    const int lane = threadIdx.x & 31;
    const long long wid = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (long long)s * words) return;
    const int i = (int)(wid / words), w = (int)(wid % words);
    const int j = w * 32 + lane;
    bool on = false;
    if (j < s) {
        float a = normalize_entry(Aa[(size_t)i * s + j], stats_a);
        if (Ab != nullptr) a = (a + normalize_entry(Ab[(size_t)i * s + j], stats_b)) / 2.f;
        on = a > tau;
    }
    unsigned word = __ballot_sync(0xffffffffu, on);
    unsigned live = __ballot_sync(0xffffffffu, j < s);
    bool keep = on && !(painted && (painted[i] || painted[j < s ? j : 0]));
    unsigned kept = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) {
        int ones = __popc(word), zeros = __popc(live) - ones;
        atomicAdd(&degree[i], (double)ones + eps * (double)zeros);
        bits[(size_t)i * words + w] = kept;
    }
}

// y = W x with W = eps 11^T + (1 - eps) B:  one warp per row, lanes stride the 32-bit words
__global__ void __launch_bounds__(256) k_bit_matvec(const uint32_t *__restrict__ bits, int s, int words, double eps,
                                                    const double *__restrict__ x, const double *__restrict__ xsum, double *__restrict__ y) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= s) return;
    double acc = 0;
    for (int w = lane; w < words; w += 32) {
        unsigned word = bits[(size_t)row * words + w];
        const double *xv = x + w * 32;
        while (word) {
            int b = __ffs(word) - 1;
            word &= word - 1;
            acc += xv[b];
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[row] = eps * xsum[0] + (1.0 - eps) * acc;
}

}  // namespace us3d

using namespace us3d;

extern "C" {

int us3d_ncut_gram(const float *f, int s, int d, float *inv_norm, float *A, uint32_t *stats, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(s > 0 && d > 0, "ncut_gram: bad shape");
    const uint32_t init[3] = {0xFFFFFFFFu, 0u, 0u};
    US3D_CUDA(cudaMemcpyAsync(stats, init, sizeof(init), cudaMemcpyHostToDevice, st));
    k_row_inv_norm<<<ceil_div(s, 8), 256, 0, st>>>(f, s, d, inv_norm);
    US3D_LAUNCH_CHECK();
    dim3 grid(ceil_div(s, 32), ceil_div(s, 32));
    k_gram<<<grid, 256, 0, st>>>(f, inv_norm, s, d, A, stats);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_ncut_threshold(const float *Aa, const float *Ab, int s, const uint32_t *stats_a, const uint32_t *stats_b, float tau,
                        double eps, const uint8_t *painted, uint32_t *bits, double *degree, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(s > 0 && Aa != nullptr, "ncut_threshold: bad arguments");
    const int words = ceil_div(s, 32);
    US3D_CUDA(cudaMemsetAsync(degree, 0, sizeof(double) * s, st));
    long long warps = (long long)s * words;
    k_threshold<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(Aa, Ab, s, words, stats_a, stats_b, tau, eps, painted, bits, degree);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_ncut_matvec(const uint32_t *bits, int s, double eps, const double *x, const double *xsum, double *y, void *stream_) {
    US3D_CHECK_ARG(s > 0, "ncut_matvec: bad shape");
    k_bit_matvec<<<ceil_div(s, 8), 256, 0, (cudaStream_t)stream_>>>(bits, s, ceil_div(s, 32), eps, x, xsum, y);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
