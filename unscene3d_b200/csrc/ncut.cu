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

// ------------------------------------------------------------------------------------------------------------------------
// Device-resident spectral step: ALL Lanczos steps of a solve in one cooperative launch (us3d_ncut_lanczos).
//
// Replaces the host-driven loop (one matvec launch + ~7 fp64 library kernels + a host decision per step, >= 512 steps per
// solve, <= 20 solves per scene) behind second_smallest_eigenvector (pseudo_masks/unscene3d_pseudo_main.py:138-146).
//
// Layout: the S elements are dealt to G = ceil(S / EPB) CTAs (EPB = 32 elements each).  A CTA keeps ITS SLICE of every
// basis vector in shared memory ([cap][EPB + 1] doubles; rows past `cap` fall back to the L2-resident global copy), so the
// two classical Gram-Schmidt passes of a step are local work: partial dots over the slice -> fp64 atomics on a global
// coefficient vector -> grid barrier -> every CTA subtracts the projection from its slice.  Per step: bit-matrix matvec of the
// CTA's rows (x = D^-1/2 q broadcast in shared memory), 3 grid barriers (after dots 1, after dots 2, after the norm).  The
// new basis row is published unnormalised together with its squared norm; readers scale on load, the owner normalises its
// global slice one phase later (no reader is left by then), so no extra barrier is needed.
namespace us3d {
namespace lz {

constexpr int THREADS = 256;
constexpr int EPB = 32;

struct Params {
    const uint32_t *bits;
    int S, words;
    double eps;
    const double *dinv;
    double *Q;
    double *alpha, *beta;
    int j0, j1, m2, cap;
    double breakdown;
    double *cA, *cB, *nrm;
    unsigned *bar;
    int *steps_done;
};

__device__ __forceinline__ double ldcg_d(const double *p) { return __ldcg(p); }

// all CTAs of the (cooperative, co-resident) grid; `gen` is the caller's count of barriers passed
__device__ __forceinline__ void grid_sync(unsigned *bar, unsigned G, unsigned &gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&bar[0], 1u) == G - 1) {
            atomicExch(&bar[0], 0u);
            __threadfence();
            atomicExch(&bar[1], gen + 1);
        } else {
            while (*((volatile unsigned *)&bar[1]) == gen) {
            }
        }
        __threadfence();
    }
    ++gen;
    __syncthreads();
}

__device__ __forceinline__ double block_sum(double v, double *red) {  // red: >= 8 doubles of shared memory; all threads get the sum
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(THREADS, 1) k_lanczos(Params p) {
    extern __shared__ __align__(16) double sm[];
    const int S = p.S, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned G = gridDim.x;
    const int e0 = blockIdx.x * EPB;
    const int ne = min(EPB, S - e0);  // valid elements of this slice
    double *x_s = sm;                          // [S]
    double *c_s = x_s + S;                     // [m2]
    double *w_s = c_s + p.m2;                  // [EPB]
    double *last_s = w_s + EPB;                // [EPB] normalised slice of the newest basis row
    double *red = last_s + EPB;                // [8][EPB] partial projections (>= 8 doubles for block_sum)
    double *Bs = red + 8 * EPB;                // [cap][EPB + 1]
    constexpr int LD = EPB + 1;
    unsigned gen = *((volatile unsigned *)&p.bar[1]);

    // basis rows 0 .. j0 + 1 are complete and normalised in global memory (host / previous launch): load the slice
    for (int idx = tid; idx < min(p.j0 + 2, p.cap) * EPB; idx += THREADS) {
        const int i = idx / EPB, e = idx - i * EPB;
        Bs[i * LD + e] = e < ne ? ldcg_d(p.Q + (size_t)i * S + e0 + e) : 0.0;
    }
    __syncthreads();

    auto basis = [&](int i, int e, int newest) -> double {  // element e of this CTA's slice of basis row i
        if (i < p.cap) return Bs[i * LD + e];
        if (i == newest) return last_s[e];
        return e < ne ? ldcg_d(p.Q + (size_t)i * S + e0 + e) : 0.0;
    };
    auto project = [&](int rows, int newest) {  // w_s[e] -= sum_i c_s[i] basis(i, e), i < rows
        const int e = tid & (EPB - 1), ig = tid / EPB;  // 8 groups of rows
        double acc = 0;
        for (int i = ig; i < rows; i += THREADS / EPB) acc += c_s[i] * basis(i, e, newest);
        red[ig * EPB + e] = acc;
        __syncthreads();
        if (tid < EPB) {
            double s = 0;
#pragma unroll
            for (int q = 0; q < THREADS / EPB; ++q) s += red[q * EPB + tid];
            w_s[tid] -= s;
        }
        __syncthreads();
    };
    auto dots = [&](int rows, int newest, double *dst) {  // dst[i] += <basis row i, w> over this slice
        for (int i = tid; i < rows; i += THREADS) {
            double s = 0;
#pragma unroll 8
            for (int e = 0; e < EPB; ++e) s += basis(i, e, newest) * w_s[e];
            atomicAdd(dst + i, s);
        }
    };

    double beta_prev = 1.0;
    int steps = p.j0;
    for (int j = p.j0; j < p.j1; ++j) {
        const int newest = j + 1, rows = j + 2;
        const double scale = j == p.j0 ? 1.0 : 1.0 / beta_prev;
        // ---- A: x = D^-1/2 q_j for ALL elements (row `newest`: normalised at j0, else as published: raw, scaled on load)
        double xs = 0;
        for (int e = tid; e < S; e += THREADS) {
            const double v = ldcg_d(p.Q + (size_t)newest * S + e) * scale * p.dinv[e];
            x_s[e] = v;
            xs += v;
        }
        if (j > p.j0 && tid < EPB) {  // this CTA's copy of the newest row was stored raw
            const double v = w_s[tid] * scale;
            last_s[tid] = v;
            if (newest < p.cap) Bs[newest * LD + tid] = v;
        } else if (j == p.j0 && tid < EPB) {
            last_s[tid] = tid < ne ? ldcg_d(p.Q + (size_t)newest * S + e0 + tid) : 0.0;
        }
        const double xsum = block_sum(xs, red);
        __syncthreads();
        // ---- matvec of this CTA's rows: y = eps * sum(x) + (1 - eps) * (B x); w = D^-1/2 y
        for (int r = warp; r < EPB; r += THREADS / 32) {
            double acc = 0;
            if (r < ne) {
                const uint32_t *brow = p.bits + (size_t)(e0 + r) * p.words;
                for (int w = lane; w < p.words; w += 32) {
                    unsigned word = brow[w];
                    const double *xv = x_s + w * 32;
                    while (word) {
                        const int b = __ffs(word) - 1;
                        word &= word - 1;
                        acc += xv[b];
                    }
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) w_s[r] = r < ne ? p.dinv[e0 + r] * (p.eps * xsum + (1.0 - p.eps) * acc) : 0.0;
        }
        if (blockIdx.x == 0)
            for (int i = tid; i < rows + 1 && i < p.m2; i += THREADS) p.cB[i] = 0.0;  // last read before barrier 3 of step j - 1
        __syncthreads();
        // ---- B: first Gram-Schmidt pass, partial coefficients
        dots(rows, newest, p.cA);
        grid_sync(p.bar, G, gen);  // #1
        // ---- C
        for (int i = tid; i < rows; i += THREADS) c_s[i] = ldcg_d(p.cA + i);
        if (j > p.j0 && tid < ne) p.Q[(size_t)newest * S + e0 + tid] = last_s[tid];  // normalise the published row (no reader left)
        if (blockIdx.x == 0 && tid == 0) p.nrm[j & 1] = 0.0;
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0) p.alpha[j] = c_s[newest];  // alpha_j = <M q_j, q_j>, taken before the orthogonalisation
        project(rows, newest);
        dots(rows, newest, p.cB);
        grid_sync(p.bar, G, gen);  // #2
        // ---- D: second pass, norm, publish the raw row
        for (int i = tid; i < rows; i += THREADS) c_s[i] = ldcg_d(p.cB + i);
        if (blockIdx.x == 0)
            for (int i = tid; i < rows + 1 && i < p.m2; i += THREADS) p.cA[i] = 0.0;  // every CTA read it before barrier 2
        __syncthreads();
        project(rows, newest);
        {
            const double v = tid < EPB ? w_s[tid] : 0.0;
            const double n2 = block_sum(v * v, red);
            if (tid == 0) atomicAdd(&p.nrm[j & 1], n2);
            if (tid < ne) p.Q[(size_t)(j + 2) * S + e0 + tid] = w_s[tid];
        }
        grid_sync(p.bar, G, gen);  // #3
        // ---- E
        const double beta = sqrt(ldcg_d(&p.nrm[j & 1]));
        if (blockIdx.x == 0 && tid == 0) p.beta[j] = beta;
        beta_prev = beta;
        steps = j + 1;
        if (beta < p.breakdown) break;  // Krylov space exhausted (the same value on every CTA)
    }
    // the last published row is still raw: normalise this CTA's slice (nobody reads it inside this launch)
    if (steps > p.j0 && beta_prev >= p.breakdown && tid < ne) p.Q[(size_t)(steps + 1) * S + e0 + tid] = w_s[tid] / beta_prev;
    if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

}  // namespace lz
}  // namespace us3d

extern "C" {

/* shared-memory capacity in basis rows for a problem of s segments and at most m steps (rows past it are read from L2) */
static int lanczos_cap(int s, int m2, size_t *smem_out) {
    const size_t fixed = sizeof(double) * ((size_t)s + m2 + 2 * lz::EPB + 8 * lz::EPB);
    const size_t budget = 200 * 1024;
    int cap = fixed + sizeof(double) * (lz::EPB + 1) < budget ? (int)((budget - fixed) / (sizeof(double) * (lz::EPB + 1))) : 0;
    if (cap > m2) cap = m2;
    *smem_out = fixed + sizeof(double) * (size_t)cap * (lz::EPB + 1);
    return cap;
}

long long us3d_ncut_lanczos_workspace_bytes(int m) { return (long long)sizeof(double) * (2 * ((long long)m + 2) + 2) + 16; }

int us3d_ncut_lanczos(const uint32_t *bits, int s, double eps, const double *dinv, double *Q, double *alpha, double *beta, int j0,
                      int j1, int m, double breakdown, void *workspace, long long workspace_bytes, int *steps_done, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(s >= 3 && m >= 1 && j0 >= 0 && j0 <= j1 && j1 <= m, "ncut_lanczos: bad step range [%d, %d) of %d", j0, j1, m);
    US3D_CHECK_ARG(workspace != nullptr && workspace_bytes >= us3d_ncut_lanczos_workspace_bytes(m), "ncut_lanczos: workspace too small");
    const int G = ceil_div(s, lz::EPB);
    US3D_CHECK_ARG(G <= num_sms(), "ncut_lanczos: %d segments need %d co-resident CTAs, the device has %d SMs", s, G, num_sms());
    lz::Params p;
    p.bits = bits; p.S = s; p.words = ceil_div(s, 32); p.eps = eps; p.dinv = dinv; p.Q = Q; p.alpha = alpha; p.beta = beta;
    p.j0 = j0; p.j1 = j1; p.m2 = m + 2; p.breakdown = breakdown;
    double *ws = (double *)workspace;
    p.cA = ws; p.cB = ws + p.m2; p.nrm = ws + 2 * (size_t)p.m2; p.bar = (unsigned *)(p.nrm + 2);
    p.steps_done = steps_done;
    size_t smem = 0;
    p.cap = lanczos_cap(s, p.m2, &smem);
    US3D_CHECK_ARG(p.cap >= 2, "ncut_lanczos: %d segments do not fit the shared-memory layout", s);
    if (j0 == 0) US3D_CUDA(cudaMemsetAsync(workspace, 0, (size_t)us3d_ncut_lanczos_workspace_bytes(m), st));
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        US3D_CUDA(cudaFuncSetAttribute(lz::k_lanczos, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        attr_done[dev] = true;
    }
    if (j0 == j1) return 0;
    void *args[] = {&p};
    US3D_CUDA(cudaLaunchCooperativeKernel((void *)lz::k_lanczos, dim3(G), dim3(lz::THREADS), args, smem, st));
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
