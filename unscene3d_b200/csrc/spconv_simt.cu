// Sparse convolution over neighbour tables — exact-fp32 SIMT path.
//
// Output-stationary: a CTA owns 64 output rows x TN output channels, walks the kernel offsets that
// have at least one present neighbour in the tile, gathers the 64 input rows of that offset into
// shared memory (zero rows for absent neighbours) and multiplies by the offset's weight slab.  No
// atomics, so results are run-to-run deterministic.  This path is the fp32 reference on the device
// and serves every layer shape (cin = 3 stem, odd channel counts); the tensor-core path in
// spconv_tc.cu takes over for the wide layers.
//
// Replaces MinkowskiConvolution / MinkowskiConvolutionTranspose forward + both gradients
// (/root/reference/models/modules/common.py:146-155, 179-188).
#include "common.cuh"

namespace us3d {

constexpr int TM = 64;   // output rows per CTA
constexpr int TK = 16;   // reduction chunk
constexpr int NT = 256;  // threads

template <int TN>
__global__ void __launch_bounds__(NT)
k_gather_conv(const float *__restrict__ x, int ldx, const int32_t *__restrict__ nbr, int n_rows, int kvol,
              const float *__restrict__ w, int cin, int cout, int transpose_w, int flip_k,
              const float *__restrict__ bias, const int32_t *__restrict__ out_rows, float *__restrict__ y, int ldy,
              int accumulate, const uint32_t *__restrict__ tile_mask, int mask_tile_rows) {
    constexpr int CT = TN / 4;       // threads along columns (4 columns each)
    constexpr int RT = NT / CT;      // threads along rows
    constexpr int RM = TM / RT;      // rows per thread
    __shared__ float As[TK][TM + 1];
    __shared__ __align__(16) float Bs[TK][TN];
    __shared__ int32_t idx_s[TM];

    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int tc = tid % CT, tr = tid / CT;
    const bool vec_ok = (cin % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const bool wvec_ok = (cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
    const bool wtvec_ok = (cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);

    float acc[RM][4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    uint32_t kmask = 0xFFFFFFFFu;
    if (tile_mask != nullptr) kmask = tile_mask[row0 / mask_tile_rows];

    for (int k = 0; k < kvol; ++k) {
        if (!((kmask >> k) & 1u)) continue;
        int valid = 0;
        __syncthreads();  // previous iteration done with idx_s / As / Bs
        if (tid < TM) {
            int j = row0 + tid;
            int r = (j < n_rows) ? nbr[(size_t)k * n_rows + j] : -1;
            idx_s[tid] = r;
            valid = r >= 0;
        }
        if (!__syncthreads_or(valid)) continue;
        const int kk = flip_k ? (kvol - 1 - k) : k;
        for (int c0 = 0; c0 < cin; c0 += TK) {
            // ---- gather A: 64 rows x 16 channels, stored [channel][row]
            {
                int r = tid >> 2, cq = (tid & 3) * 4;
                int src = idx_s[r];
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (src >= 0) {
                    const float *p = x + (size_t)src * ldx + c0 + cq;
                    if (vec_ok && c0 + cq + 3 < cin) {
                        float4 t = *reinterpret_cast<const float4 *>(p);
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (c0 + cq + i < cin) v[i] = p[i];
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) As[cq + i][r] = v[i];
            }
            // ---- weights B: 16 channels x TN outputs
            if (!transpose_w) {
                for (int e = tid; e < TK * CT; e += NT) {
                    int c = e / CT, nq = (e % CT) * 4;
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (c0 + c < cin) {
                        const float *p = w + ((size_t)kk * cin + c0 + c) * cout + n0 + nq;
                        if (wvec_ok && n0 + nq + 3 < cout) {
                            float4 t = *reinterpret_cast<const float4 *>(p);
                            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (n0 + nq + i < cout) v[i] = p[i];
                        }
                    }
                    *reinterpret_cast<float4 *>(&Bs[c][nq]) = make_float4(v[0], v[1], v[2], v[3]);
                }
            } else {
                // W is [kvol, cout, cin] physically; logical Wk[c][n] = W[kk][n][c]
                for (int e = tid; e < TN * (TK / 4); e += NT) {
                    int n = e / (TK / 4), cq = (e % (TK / 4)) * 4;
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (n0 + n < cout) {
                        const float *p = w + ((size_t)kk * cout + n0 + n) * cin + c0 + cq;
                        if (wtvec_ok && c0 + cq + 3 < cin) {
                            float4 t = *reinterpret_cast<const float4 *>(p);
                            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (c0 + cq + i < cin) v[i] = p[i];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) Bs[cq + i][n] = v[i];
                }
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < TK; ++c) {
                float4 b = *reinterpret_cast<const float4 *>(&Bs[c][tc * 4]);
#pragma unroll
                for (int i = 0; i < RM; ++i) {
                    float a = As[c][tr * RM + i];
                    acc[i][0] = fmaf(a, b.x, acc[i][0]);
                    acc[i][1] = fmaf(a, b.y, acc[i][1]);
                    acc[i][2] = fmaf(a, b.z, acc[i][2]);
                    acc[i][3] = fmaf(a, b.w, acc[i][3]);
                }
            }
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < RM; ++i) {
        int j = row0 + tr * RM + i;
        if (j >= n_rows) continue;
        int orow = out_rows ? out_rows[j] : j;
        float *py = y + (size_t)orow * ldy + n0 + tc * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int n = n0 + tc * 4 + c;
            if (n >= cout) continue;
            float v = acc[i][c] + (bias ? bias[n] : 0.f);
            py[c] = accumulate ? py[c] + v : v;
        }
    }
}

// dW[k][ci][co] += sum_j X[nbr[k][j]][ci] * dY[orow(j)][co]   — CTA = (k, 64x64 tile of dW, row split)
__global__ void __launch_bounds__(NT)
k_wgrad(const float *__restrict__ x, int ldx, const int32_t *__restrict__ nbr, int n_rows, int kvol,
        const float *__restrict__ dy, int ldy, const int32_t *__restrict__ out_rows, float *__restrict__ dw, int cin,
        int cout, int rows_per_split, int co_tiles) {
    __shared__ __align__(16) float As[TK][64];
    __shared__ __align__(16) float Bs[TK][64];
    __shared__ int32_t idx_s[TK];
    __shared__ int32_t orow_s[TK];
    const int tid = threadIdx.x;
    const int k = blockIdx.x;
    const int ci0 = (blockIdx.y / co_tiles) * 64, co0 = (blockIdx.y % co_tiles) * 64;
    const int r_begin = blockIdx.z * rows_per_split;
    const int r_end = min(n_rows, r_begin + rows_per_split);
    const int tc = tid % 16, tr = tid / 16;
    const bool xv = (cin % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const bool yv = (cout % 4 == 0) && (ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int r0 = r_begin; r0 < r_end; r0 += TK) {
        int valid = 0;
        __syncthreads();
        if (tid < TK) {
            int j = r0 + tid;
            int src = (j < r_end) ? nbr[(size_t)k * n_rows + j] : -1;
            idx_s[tid] = src;
            orow_s[tid] = (src >= 0) ? (out_rows ? out_rows[j] : j) : -1;
            valid = src >= 0;
        }
        if (!__syncthreads_or(valid)) continue;
        {
            int r = tid / 16, cq = (tid % 16) * 4;
            int src = idx_s[r];
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (src >= 0) {
                const float *p = x + (size_t)src * ldx + ci0 + cq;
                if (xv && ci0 + cq + 3 < cin) a = *reinterpret_cast<const float4 *>(p);
                else {
                    if (ci0 + cq + 0 < cin) a.x = p[0];
                    if (ci0 + cq + 1 < cin) a.y = p[1];
                    if (ci0 + cq + 2 < cin) a.z = p[2];
                    if (ci0 + cq + 3 < cin) a.w = p[3];
                }
                const float *q = dy + (size_t)orow_s[r] * ldy + co0 + cq;
                if (yv && co0 + cq + 3 < cout) b = *reinterpret_cast<const float4 *>(q);
                else {
                    if (co0 + cq + 0 < cout) b.x = q[0];
                    if (co0 + cq + 1 < cout) b.y = q[1];
                    if (co0 + cq + 2 < cout) b.z = q[2];
                    if (co0 + cq + 3 < cout) b.w = q[3];
                }
            }
            *reinterpret_cast<float4 *>(&As[r][cq]) = a;
            *reinterpret_cast<float4 *>(&Bs[r][cq]) = b;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < TK; ++r) {
            float4 a = *reinterpret_cast<const float4 *>(&As[r][tr * 4]);
            float4 b = *reinterpret_cast<const float4 *>(&Bs[r][tc * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int ci = ci0 + tr * 4 + i;
        if (ci >= cin) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = co0 + tc * 4 + j;
            if (co >= cout) continue;
            if (acc[i][j] != 0.f) atomicAdd(&dw[((size_t)k * cin + ci) * cout + co], acc[i][j]);
        }
    }
}

}  // namespace us3d

using namespace us3d;

extern "C" {

int us3d_spconv_gather(const float *x, int ldx, const int32_t *nbr, int n_rows, int kvol, const float *w, int cin,
                       int cout, int transpose_w, int flip_k, const float *bias, const int32_t *out_rows, float *y,
                       int ldy, int accumulate, const uint32_t *tile_mask, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "spconv_gather: kvol %d out of range", kvol);
    US3D_CHECK_ARG(cin > 0 && cout > 0 && ldx >= cin && ldy >= cout, "spconv_gather: bad channel counts / leading dims");
    if (n_rows == 0) return 0;
    const int mask_tile_rows = 128;  // tile masks are always built for 128-row tiles
    ProfScope prof(st, 0, n_rows, n_rows, kvol, cin, cout);
    if (cout % 64 == 0 || cout > 96) {
        dim3 grid(ceil_div(n_rows, TM), ceil_div(cout, 64));
        k_gather_conv<64><<<grid, NT, 0, st>>>(x, ldx, nbr, n_rows, kvol, w, cin, cout, transpose_w, flip_k, bias, out_rows,
                                               y, ldy, accumulate, tile_mask, mask_tile_rows);
    } else {
        dim3 grid(ceil_div(n_rows, TM), ceil_div(cout, 32));
        k_gather_conv<32><<<grid, NT, 0, st>>>(x, ldx, nbr, n_rows, kvol, w, cin, cout, transpose_w, flip_k, bias, out_rows,
                                               y, ldy, accumulate, tile_mask, mask_tile_rows);
    }
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_spconv_wgrad(const float *x, int ldx, const int32_t *nbr, int n_rows, int kvol, const float *dy, int ldy,
                      const int32_t *out_rows, float *dw, int cin, int cout, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "spconv_wgrad: kvol %d out of range", kvol);
    US3D_CHECK_ARG(cin > 0 && cout > 0 && ldx >= cin && ldy >= cout, "spconv_wgrad: bad channel counts / leading dims");
    if (n_rows == 0) return 0;
    int ci_tiles = ceil_div(cin, 64), co_tiles = ceil_div(cout, 64);
    int tiles = kvol * ci_tiles * co_tiles;
    int splits = ceil_div(4 * num_sms(), tiles);
    int max_splits = ceil_div(n_rows, 256);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int rows_per_split = ceil_div(ceil_div(n_rows, splits), TK) * TK;
    splits = ceil_div(n_rows, rows_per_split);
    dim3 grid(kvol, ci_tiles * co_tiles, splits);
    ProfScope prof(st, 1, n_rows, n_rows, kvol, cin, cout);
    k_wgrad<<<grid, NT, 0, st>>>(x, ldx, nbr, n_rows, kvol, dy, ldy, out_rows, dw, cin, cout, rows_per_split, co_tiles);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
