// Masked multi-head cross-attention of the Mask3D decoder: the attention core of
// models/mask3d.py:355-365 — nn.MultiheadAttention(d=128, h=8) called from CrossAttentionLayer
// (models/mask3d.py:561-651) with a boolean memory_mask (True = key hidden from the query).
//
// Shapes of the path: Q = 100 queries, K <= 12 800 sampled voxels per scene, head_dim = 16, B <= 4 scenes: 0.33 GFLOP per
// round, so the contraction is not the cost — the reference's route is: repeat_interleave of the [B,K,Q] mask to
// [B*h,Q,K], conversion to a float -inf mask, a materialised [B*h,Q,K] score tensor, softmax, a second batched
// GEMM, and the same again in backward (~160 MB of fp32 per round at C3).  Here nothing of size Q x K is ever written:
//   forward   k_xattn_fwd      split-K flash attention, thread = query, keys/values of a 128-key chunk in shared memory,
//                              running (max, sum, acc[hd]) in registers, partials per chunk
//             k_xattn_combine  merges the chunk partials, writes O and the row log-sum-exp
//   backward  k_xattn_bwd_kv   thread = key: dK, dV of a key are complete in one thread (no atomics, deterministic)
//             k_xattn_bwd_q    thread = query: partial dQ per chunk;  k_xattn_reduce_dq sums the chunks in fixed order
// The mask is read where it lies through (batch, head, query, key) strides, so both the decoder's own [B,K,Q] layout and
// the reference's [B*h,Q,K] layout are consumed without a copy.  fp32 throughout (SIMT): K = 16 contractions do not fill
// a tensor-core tile and the kernel is latency/issue bound, not FLOP bound.
#include "common.cuh"

namespace us3d {
namespace attn {

constexpr int kThreads = 128;  // queries (fwd, bwd_q) or keys (bwd_kv) per CTA
constexpr int kChunk = 128;    // keys per CTA

struct MaskView {
    const uint8_t *p;  // may be NULL = nothing masked
    long long sb, sh, sq, sk;
};

__device__ __forceinline__ unsigned allowed_bits(const MaskView &mk, long long base, long long step, int count) {
    // bit i set <=> element base + i*step is visible (mask byte == 0); count <= 32
    if (mk.p == nullptr) return count >= 32 ? 0xffffffffu : ((1u << count) - 1u);
    unsigned bits = 0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i)
        if (i < count && mk.p[base + (long long)i * step] == 0) bits |= 1u << i;
    return bits;
}

template <int HD>
__device__ __forceinline__ void load_rows(float (*dst)[HD], const float *src, long long row_stride, int row0, int rows, int total) {
    // dst[r][:] = src[(row0 + r) * row_stride + 0..HD), zero beyond `total` rows
    constexpr int V = HD / 4;
    for (int e = threadIdx.x; e < rows * V; e += blockDim.x) {
        int r = e / V, c = e % V;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < total) v = *reinterpret_cast<const float4 *>(src + (long long)(row0 + r) * row_stride + c * 4);
        *reinterpret_cast<float4 *>(&dst[r][c * 4]) = v;
    }
}

// ------------------------------------------------------------------------------------------------ forward
// ws: [B][H][nsplit][Q][HD + 2]  (running max, running sum, acc[HD])
template <int HD>
__global__ void __launch_bounds__(kThreads)
k_xattn_fwd(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v, MaskView mk, int B, int H,
            int Q, int K, float scale, int nsplit, float *__restrict__ ws) {
    __shared__ __align__(16) float ks[kChunk][HD];
    __shared__ __align__(16) float vs[kChunk][HD];
    const int split = blockIdx.x % nsplit, qblk = blockIdx.x / nsplit, h = blockIdx.y, b = blockIdx.z;
    const int E = H * HD;
    const long long row = (long long)B * E;
    const int key0 = split * kChunk, kc = min(kChunk, K - key0);
    load_rows<HD>(ks, k + (long long)b * E + h * HD, row, key0, kChunk, K);
    load_rows<HD>(vs, v + (long long)b * E + h * HD, row, key0, kChunk, K);
    __syncthreads();
    const int qi = qblk * kThreads + threadIdx.x;
    if (qi >= Q) return;
    float qr[HD], acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) qr[c] = q[(long long)qi * row + (long long)b * E + h * HD + c] * scale, acc[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    const long long mbase = (long long)b * mk.sb + (long long)h * mk.sh + (long long)qi * mk.sq;
    for (int j0 = 0; j0 < kc; j0 += 32) {
        const unsigned allowed = allowed_bits(mk, mbase + (long long)(key0 + j0) * mk.sk, mk.sk, min(32, kc - j0));
#pragma unroll 4
        for (int jj = 0; jj < 32; ++jj) {
            if (!(allowed >> jj & 1u)) continue;
            const int j = j0 + jj;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) s = fmaf(qr[c], ks[j][c], s);
            if (s > m) {
                const float r = expf(m - s);
                l *= r;
#pragma unroll
                for (int c = 0; c < HD; ++c) acc[c] *= r;
                m = s;
            }
            const float p = expf(s - m);
            l += p;
#pragma unroll
            for (int c = 0; c < HD; ++c) acc[c] = fmaf(p, vs[j][c], acc[c]);
        }
    }
    float *o = ws + ((((long long)b * H + h) * nsplit + split) * Q + qi) * (HD + 2);
    o[0] = m, o[1] = l;
#pragma unroll
    for (int c = 0; c < HD; ++c) o[2 + c] = acc[c];
}

template <int HD>
__global__ void __launch_bounds__(kThreads)
k_xattn_combine(const float *__restrict__ ws, int B, int H, int Q, int nsplit, float *__restrict__ out, float *__restrict__ lse) {
    const int qi = blockIdx.x * kThreads + threadIdx.x, h = blockIdx.y, b = blockIdx.z;
    if (qi >= Q) return;
    const float *p = ws + (((long long)b * H + h) * nsplit * Q + qi) * (HD + 2);
    const long long step = (long long)Q * (HD + 2);
    float M = -INFINITY;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, p[s * step]);
    float l = 0.f, acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        const float ms = p[s * step];
        if (ms == -INFINITY) continue;  // chunk fully hidden from this query
        const float w = expf(ms - M);
        l = fmaf(p[s * step + 1], w, l);
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = fmaf(p[s * step + 2 + c], w, acc[c]);
    }
    const float inv = 1.f / l;  // l == 0 (every key hidden): NaN, as torch's softmax over an all -inf row
    float *o = out + (long long)qi * B * H * HD + (long long)b * H * HD + h * HD;
#pragma unroll
    for (int c = 0; c < HD; ++c) o[c] = acc[c] * inv;
    lse[((long long)b * H + h) * Q + qi] = M + logf(l);
}

// ------------------------------------------------------------------------------------------------ backward
// dK, dV: one thread per key, queries streamed through shared memory in tiles of 128
template <int HD>
__global__ void __launch_bounds__(kThreads)
k_xattn_bwd_kv(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v, MaskView mk,
               const float *__restrict__ out, const float *__restrict__ lse, const float *__restrict__ dout, int B, int H, int Q,
               int K, float scale, float *__restrict__ dk, float *__restrict__ dv) {
    __shared__ __align__(16) float qs[kThreads][HD];
    __shared__ __align__(16) float dos[kThreads][HD];
    __shared__ float Ls[kThreads], Ds[kThreads];
    const int h = blockIdx.y, b = blockIdx.z, E = H * HD;
    const long long row = (long long)B * E, col = (long long)b * E + h * HD;
    const int key = blockIdx.x * kChunk + threadIdx.x;
    const bool live = key < K;
    float kr[HD], vr[HD], dkr[HD], dvr[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) {
        kr[c] = live ? k[(long long)key * row + col + c] : 0.f;
        vr[c] = live ? v[(long long)key * row + col + c] : 0.f;
        dkr[c] = dvr[c] = 0.f;
    }
    const long long mbase = (long long)b * mk.sb + (long long)h * mk.sh + (long long)key * mk.sk;
    for (int q0 = 0; q0 < Q; q0 += kThreads) {
        __syncthreads();
        load_rows<HD>(qs, q + col, row, q0, kThreads, Q);
        load_rows<HD>(dos, dout + col, row, q0, kThreads, Q);
        {
            const int qi = q0 + threadIdx.x;
            float d = 0.f;
            if (qi < Q) {
#pragma unroll
                for (int c = 0; c < HD; ++c) d = fmaf(dout[(long long)qi * row + col + c], out[(long long)qi * row + col + c], d);
                Ls[threadIdx.x] = lse[((long long)b * H + h) * Q + qi];
            }
            Ds[threadIdx.x] = d;
        }
        __syncthreads();
        const int qn = min(kThreads, Q - q0);
        if (!live) continue;
        for (int t0 = 0; t0 < qn; t0 += 32) {
            const unsigned allowed = allowed_bits(mk, mbase + (long long)(q0 + t0) * mk.sq, mk.sq, min(32, qn - t0));
#pragma unroll 4
            for (int tt = 0; tt < 32; ++tt) {
                if (!(allowed >> tt & 1u)) continue;
                const int t = t0 + tt;
                float s = 0.f, dp = 0.f;
#pragma unroll
                for (int c = 0; c < HD; ++c) s = fmaf(qs[t][c], kr[c], s), dp = fmaf(dos[t][c], vr[c], dp);
                const float p = expf(s * scale - Ls[t]);
                const float ds = p * (dp - Ds[t]) * scale;
#pragma unroll
                for (int c = 0; c < HD; ++c) dvr[c] = fmaf(p, dos[t][c], dvr[c]), dkr[c] = fmaf(ds, qs[t][c], dkr[c]);
            }
        }
    }
    if (live) {
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
            *reinterpret_cast<float4 *>(dk + (long long)key * row + col + c) = make_float4(dkr[c], dkr[c + 1], dkr[c + 2], dkr[c + 3]);
            *reinterpret_cast<float4 *>(dv + (long long)key * row + col + c) = make_float4(dvr[c], dvr[c + 1], dvr[c + 2], dvr[c + 3]);
        }
    }
}

// partial dQ per key chunk: ws [B][H][nsplit][Q][HD]
template <int HD>
__global__ void __launch_bounds__(kThreads)
k_xattn_bwd_q(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v, MaskView mk,
              const float *__restrict__ out, const float *__restrict__ lse, const float *__restrict__ dout, int B, int H, int Q,
              int K, float scale, int nsplit, float *__restrict__ ws) {
    __shared__ __align__(16) float ks[kChunk][HD];
    __shared__ __align__(16) float vs[kChunk][HD];
    const int split = blockIdx.x % nsplit, qblk = blockIdx.x / nsplit, h = blockIdx.y, b = blockIdx.z;
    const int E = H * HD;
    const long long row = (long long)B * E, col = (long long)b * E + h * HD;
    const int key0 = split * kChunk, kc = min(kChunk, K - key0);
    load_rows<HD>(ks, k + col, row, key0, kChunk, K);
    load_rows<HD>(vs, v + col, row, key0, kChunk, K);
    __syncthreads();
    const int qi = qblk * kThreads + threadIdx.x;
    if (qi >= Q) return;
    float qr[HD], dor[HD], dq[HD];
    float D = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) {
        qr[c] = q[(long long)qi * row + col + c] * scale;
        dor[c] = dout[(long long)qi * row + col + c];
        D = fmaf(dor[c], out[(long long)qi * row + col + c], D);
        dq[c] = 0.f;
    }
    const float L = lse[((long long)b * H + h) * Q + qi];
    const long long mbase = (long long)b * mk.sb + (long long)h * mk.sh + (long long)qi * mk.sq;
    for (int j0 = 0; j0 < kc; j0 += 32) {
        const unsigned allowed = allowed_bits(mk, mbase + (long long)(key0 + j0) * mk.sk, mk.sk, min(32, kc - j0));
#pragma unroll 4
        for (int jj = 0; jj < 32; ++jj) {
            if (!(allowed >> jj & 1u)) continue;
            const int j = j0 + jj;
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) s = fmaf(qr[c], ks[j][c], s), dp = fmaf(dor[c], vs[j][c], dp);
            const float ds = expf(s - L) * (dp - D);
#pragma unroll
            for (int c = 0; c < HD; ++c) dq[c] = fmaf(ds, ks[j][c], dq[c]);
        }
    }
    float *o = ws + ((((long long)b * H + h) * nsplit + split) * Q + qi) * HD;
#pragma unroll
    for (int c = 0; c < HD; ++c) o[c] = dq[c];
}

template <int HD>
__global__ void __launch_bounds__(256)
k_xattn_reduce_dq(const float *__restrict__ ws, int B, int H, int Q, int nsplit, float scale, float *__restrict__ dq) {
    const long long total = (long long)B * H * Q * HD;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % HD);
        const int qi = (int)(e / HD % Q);
        const int h = (int)(e / ((long long)HD * Q) % H);
        const int b = (int)(e / ((long long)HD * Q * H));
        const float *p = ws + (((long long)b * H + h) * nsplit * Q + qi) * HD + c;
        float s = 0.f;
        for (int i = 0; i < nsplit; ++i) s += p[(long long)i * Q * HD];
        dq[(long long)qi * B * H * HD + (long long)b * H * HD + h * HD + c] = s * scale;
    }
}

template <int HD>
int fwd(const float *q, const float *k, const float *v, MaskView mk, int B, int H, int Q, int K, float scale, float *ws, float *out,
        float *lse, cudaStream_t st) {
    const int nsplit = ceil_div(K, kChunk), qblks = ceil_div(Q, kThreads);
    k_xattn_fwd<HD><<<dim3(nsplit * qblks, H, B), kThreads, 0, st>>>(q, k, v, mk, B, H, Q, K, scale, nsplit, ws);
    US3D_LAUNCH_CHECK();
    k_xattn_combine<HD><<<dim3(qblks, H, B), kThreads, 0, st>>>(ws, B, H, Q, nsplit, out, lse);
    US3D_LAUNCH_CHECK();
    return 0;
}

template <int HD>
int bwd(const float *q, const float *k, const float *v, MaskView mk, const float *out, const float *lse, const float *dout, int B,
        int H, int Q, int K, float scale, float *ws, float *dq, float *dk, float *dv, cudaStream_t st) {
    const int nsplit = ceil_div(K, kChunk), qblks = ceil_div(Q, kThreads);
    k_xattn_bwd_kv<HD><<<dim3(nsplit, H, B), kThreads, 0, st>>>(q, k, v, mk, out, lse, dout, B, H, Q, K, scale, dk, dv);
    US3D_LAUNCH_CHECK();
    k_xattn_bwd_q<HD><<<dim3(nsplit * qblks, H, B), kThreads, 0, st>>>(q, k, v, mk, out, lse, dout, B, H, Q, K, scale, nsplit, ws);
    US3D_LAUNCH_CHECK();
    const long long total = (long long)B * H * Q * HD;
    k_xattn_reduce_dq<HD><<<ceil_div(total, 256), 256, 0, st>>>(ws, B, H, Q, nsplit, scale, dq);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // namespace attn
}  // namespace us3d

using namespace us3d;

extern "C" {

long long us3d_xattn_workspace_bytes(int b, int h, int q, int k, int head_dim) {
    if (b <= 0 || h <= 0 || q <= 0 || k <= 0 || head_dim <= 0) return 0;
    return (long long)b * h * ceil_div(k, attn::kChunk) * q * (head_dim + 2) * (long long)sizeof(float);
}

int us3d_xattn_fwd(const float *q, const float *k, const float *v, const uint8_t *mask, long long mask_sb, long long mask_sh,
                   long long mask_sq, long long mask_sk, int b, int h, int nq, int nk, int head_dim, float scale, float *ws,
                   float *out, float *lse, void *stream_) {
    US3D_CHECK_ARG(b > 0 && h > 0 && nq > 0 && nk > 0, "xattn_fwd: bad shape");
    US3D_CHECK_ARG(head_dim == 16 || head_dim == 32, "xattn_fwd: head_dim %d not supported (16 or 32)", head_dim);
    US3D_CHECK_ARG(h <= 65535 && b <= 65535, "xattn_fwd: heads/batch exceed the grid limits");
    attn::MaskView mk{mask, mask_sb, mask_sh, mask_sq, mask_sk};
    cudaStream_t st = (cudaStream_t)stream_;
    return head_dim == 16 ? attn::fwd<16>(q, k, v, mk, b, h, nq, nk, scale, ws, out, lse, st)
                          : attn::fwd<32>(q, k, v, mk, b, h, nq, nk, scale, ws, out, lse, st);
}

int us3d_xattn_bwd(const float *q, const float *k, const float *v, const uint8_t *mask, long long mask_sb, long long mask_sh,
                   long long mask_sq, long long mask_sk, const float *out, const float *lse, const float *dout, int b, int h, int nq,
                   int nk, int head_dim, float scale, float *ws, float *dq, float *dk, float *dv, void *stream_) {
    US3D_CHECK_ARG(b > 0 && h > 0 && nq > 0 && nk > 0, "xattn_bwd: bad shape");
    US3D_CHECK_ARG(head_dim == 16 || head_dim == 32, "xattn_bwd: head_dim %d not supported (16 or 32)", head_dim);
    US3D_CHECK_ARG(h <= 65535 && b <= 65535, "xattn_bwd: heads/batch exceed the grid limits");
    attn::MaskView mk{mask, mask_sb, mask_sh, mask_sq, mask_sk};
    cudaStream_t st = (cudaStream_t)stream_;
    return head_dim == 16 ? attn::bwd<16>(q, k, v, mk, out, lse, dout, b, h, nq, nk, scale, ws, dq, dk, dv, st)
                          : attn::bwd<32>(q, k, v, mk, out, lse, dout, b, h, nq, nk, scale, ws, dq, dk, dv, st);
}

}  // extern "C"
