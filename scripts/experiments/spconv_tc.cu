// Sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Output-stationary implicit GEMM over a neighbour table (same contract as us3d_spconv_gather):
//     Y[orow(j)] = sum_k X[nbr[k, j]] · W'[k]            one CTA = 128 output rows x all Cout columns
//
//   warps 0-3  gather producers, later the epilogue.  For every (active offset k, 64-channel chunk) they
//              read the 128 neighbour rows (fp32, 256 B contiguous per row and chunk, 8 lanes per row so
//              every 32 B sector is fully used; absent neighbours become zero rows), split each value
//              into bf16 hi (+ bf16 lo = x - hi when PASSES == 3) and store the 128 x 64 tile into shared
//              memory in the canonical K-major SWIZZLE_128B UMMA layout.
//   warp 5     streams the pre-packed weight slab of (k, chunk) with one cp.async.bulk per plane; the
//              slab is stored in global memory already in its swizzled shared-memory image
//              (us3d_spconv_pack_weights), so no tensor map is needed.
//   warp 4     one elected lane issues tcgen05.mma (M=128, N=Cout, K=16, kind::f16, fp32 accumulate in TMEM)
//              for every 16-channel step: hi·hi, and for PASSES == 3 also lo·hi and hi·lo — the three-term
//              bf16 split reproduces fp32 products to ~2^-17 relative; tcgen05.commit releases the stage.
//   epilogue   tcgen05.ld 32 lanes x 32 columns per warp -> registers -> (+bias, +=) -> fp32 rows in HBM.
//
// Full/empty mbarriers form a STAGES-deep ring between producers and the MMA warp; offsets whose bit is
// clear in the tile mask (no neighbour in the whole tile) are skipped by every role.
#include <cuda_bf16.h>

#include "common.cuh"

namespace us3d {
namespace tc {

constexpr int M = 128;        // output rows per CTA (UMMA M)
constexpr int KC = 64;        // channels per stage (one 128-byte swizzle row of bf16)
constexpr int PROD_WARPS = 8;  // gather producers; the same warps run the epilogue
constexpr int THREADS = (PROD_WARPS + 2) * 32;  // + MMA warp + weight warp
constexpr int MAX_STAGES = 4;
constexpr int A_BYTES = M * 128;  // one plane of the A tile

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU — trap after ~2 s instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("us3d spconv_tc: mbarrier wait timed out (role %d, block %d, thread %d, parity %u)\n", who, blockIdx.x,
                   threadIdx.x, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                   // leading byte offset (ignored for swizzled K-major), canonical 1
    d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                   // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}

struct Params {
    const float *x;
    int ldx;
    const int32_t *nbr;
    int n_rows, kvol;
    const uint8_t *wpack;  // [kvol][nchunks][PASSES==3 ? 2 : 1][cout][128 B], swizzled
    int cin, cout, nchunks;
    const float *bias;
    const int32_t *out_rows;
    float *y;
    int ldy, accumulate;
    const uint32_t *tile_mask;
    int stages, b_bytes, stage_bytes, tmem_cols;
};

template <int PASSES>
__global__ void __launch_bounds__(THREADS, 2) k_spconv_tc(Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // barriers live in static shared memory, tiles in the (1024-byte aligned) dynamic part
    __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tmem;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile0 = blockIdx.x * M;
    constexpr int NPL = PASSES == 3 ? 2 : 1;  // planes per operand
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);

    uint32_t kmask = p.kvol >= 32 ? 0xFFFFFFFFu : ((1u << p.kvol) - 1u);
    if (p.tile_mask != nullptr) kmask &= p.tile_mask[blockIdx.x];
    const int n_active = __popc(kmask);
    const int n_items = n_active * p.nchunks;

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), PROD_WARPS + 1);  // producer warps + the weight warp's expect_tx arrive
            mbar_init(smem_u32(&bar_empty[s]), 1);     // tcgen05.commit
        }
        mbar_init(smem_u32(&bar_tmem), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < PROD_WARPS) {
        // ------------------------------------------------------------------ gather producers
        const int grp = tid & 7;        // 8-channel group inside the 64-channel chunk
        const int rbase = tid >> 3;     // rows rbase + 32 i
        constexpr int RPT = M * 8 / (PROD_WARPS * 32);  // rows per thread (4)
        int item = 0;
        for (int k = 0; k < p.kvol; ++k) {
            if (!((kmask >> k) & 1u)) continue;
            int idx[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                int j = tile0 + rbase + (M / RPT) * i;
                idx[i] = j < p.n_rows ? __ldg(p.nbr + (size_t)k * p.n_rows + j) : -1;
            }
            for (int c = 0; c < p.nchunks; ++c, ++item) {
                const int s = item % p.stages;
                const uint32_t par = (item / p.stages) & 1;
                mbar_wait(smem_u32(&bar_empty[s]), par ^ 1, 0);
                uint8_t *a_hi = smem + (size_t)s * p.stage_bytes;
                uint8_t *a_lo = a_hi + A_BYTES;
                const int c0 = c * KC + grp * 8;
                const bool in_range = c0 < p.cin;  // partial last chunk: upper groups are never read by the MMA
                float4 v[RPT][2];
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    if (idx[i] >= 0 && in_range) {
                        const float4 *src = reinterpret_cast<const float4 *>(p.x + (size_t)idx[i] * p.ldx + c0);
                        v[i][0] = __ldg(src);
                        v[i][1] = __ldg(src + 1);
                    } else {
                        v[i][0] = v[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (in_range) {
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const int r = rbase + (M / RPT) * i;
                        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((grp ^ (r & 7)) << 4);
                        const float f[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
                        uint4 hi;
                        hi.x = pack_bf16(f[0], f[1]);
                        hi.y = pack_bf16(f[2], f[3]);
                        hi.z = pack_bf16(f[4], f[5]);
                        hi.w = pack_bf16(f[6], f[7]);
                        *reinterpret_cast<uint4 *>(a_hi + off) = hi;
                        if (PASSES == 3) {
                            float l[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) l[e] = f[e] - __bfloat162float(__float2bfloat16_rn(f[e]));
                            uint4 lo;
                            lo.x = pack_bf16(l[0], l[1]);
                            lo.y = pack_bf16(l[2], l[3]);
                            lo.z = pack_bf16(l[4], l[5]);
                            lo.w = pack_bf16(l[6], l[7]);
                            *reinterpret_cast<uint4 *>(a_lo + off) = lo;
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
            }
        }
        // ------------------------------------------------------------------ epilogue
        // warp w reads TMEM lanes 32 (w % 4) .. +31 (its row quarter); the two warps sharing a quarter
        // alternate 16-column batches
        const int quarter = warp & 3, half = warp >> 2;
        const int r = quarter * 32 + lane;
        const int j = tile0 + r;
        const bool row_ok = j < p.n_rows;
        float *yrow = nullptr;
        if (row_ok) yrow = p.y + (size_t)(p.out_rows ? p.out_rows[j] : j) * p.ldy;
        const bool vec = (p.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        if (n_items > 0) {
            mbar_wait(smem_u32(&bar_tmem), 0, 1);
            tc_fence_after();
        }
        for (int col = half * 16; col < p.cout; col += 16 * (PROD_WARPS / 4)) {
            float acc[16];
            if (n_items > 0) {
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col, acc);
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[e] = 0.f;
            }
            if (!row_ok) continue;
            if (p.bias)
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[e] += __ldg(p.bias + col + e);
            if (vec) {
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    float4 o = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
                    float4 *dst = reinterpret_cast<float4 *>(yrow + col + e);
                    if (p.accumulate) {
                        float4 old = *dst;
                        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                    }
                    *dst = o;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) yrow[col + e] = p.accumulate ? yrow[col + e] + acc[e] : acc[e];
            }
        }
    } else if (warp == PROD_WARPS) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.cout);
            for (int item = 0; item < n_items; ++item) {
                const int s = item % p.stages;
                const uint32_t par = (item / p.stages) & 1;
                const int c = item % p.nchunks;
                const int ksteps = min(KC, p.cin - c * KC) / 16;
                mbar_wait(smem_u32(&bar_full[s]), par, 2);
                tc_fence_after();
                const uint32_t a_hi = smem_base + (uint32_t)s * p.stage_bytes;
                const uint32_t a_lo = a_hi + A_BYTES;
                const uint32_t b_hi = a_hi + NPL * A_BYTES;
                const uint32_t b_lo = b_hi + p.b_bytes;
                const uint64_t da_hi = make_desc_sw128(a_hi), da_lo = make_desc_sw128(a_lo);
                const uint64_t db_hi = make_desc_sw128(b_hi), db_lo = make_desc_sw128(b_lo);
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint64_t adv = (uint64_t)(kk * 2);  // 16 bf16 = 32 B = 2 x 16-byte units
                    umma(tmem_base, da_hi + adv, db_hi + adv, idesc, (item | kk) != 0);
                    if (PASSES == 3) {
                        umma(tmem_base, da_lo + adv, db_hi + adv, idesc, 1);
                        umma(tmem_base, da_hi + adv, db_lo + adv, idesc, 1);
                    }
                }
                umma_commit(smem_u32(&bar_empty[s]));
            }
            if (n_items > 0) umma_commit(smem_u32(&bar_tmem));
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ weight slabs
        if (lane == 0) {
            int item = 0;
            for (int k = 0; k < p.kvol; ++k) {
                if (!((kmask >> k) & 1u)) continue;
                for (int c = 0; c < p.nchunks; ++c, ++item) {
                    const int s = item % p.stages;
                    const uint32_t par = (item / p.stages) & 1;
                    mbar_wait(smem_u32(&bar_empty[s]), par ^ 1, 3);
                    const uint32_t bar = smem_u32(&bar_full[s]);
                    const uint32_t dst = smem_base + (uint32_t)s * p.stage_bytes + NPL * A_BYTES;
                    const uint8_t *src = p.wpack + ((size_t)k * p.nchunks + c) * NPL * (size_t)p.b_bytes;
                    mbar_arrive_expect_tx(bar, (uint32_t)(NPL * p.b_bytes));
                    bulk_g2s(dst, src, (uint32_t)p.b_bytes, bar);
                    if (PASSES == 3) bulk_g2s(dst + p.b_bytes, src + p.b_bytes, (uint32_t)p.b_bytes, bar);
                }
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// Weight pre-pack: fp32 W -> bf16 hi (+lo) planes in the swizzled K-major shared-memory image.
//   logical B_k[n][kk] = W[kq][kk][n] (forward) or W[kq][n][kk] (transpose: dgrad), kq = flip ? kvol-1-k : k
__global__ void __launch_bounds__(256) k_pack_weights(const float *__restrict__ w, int kvol, int kdim, int ndim, int w_cin,
                                                      int w_cout, int transpose, int flip, int planes, uint8_t *__restrict__ out) {
    const int nchunks = (kdim + KC - 1) / KC;
    const long long total = (long long)kvol * nchunks * ndim * 8;
    const size_t slab = (size_t)ndim * 128;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(e % ndim);
        long long t = e / ndim;
        const int g = (int)(t % 8);
        t /= 8;
        const int c = (int)(t % nchunks);
        const int k = (int)(t / nchunks);
        const int kq = flip ? kvol - 1 - k : k;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kk = c * KC + g * 8 + i;
            float v = 0.f;
            if (kk < kdim) v = transpose ? w[((size_t)kq * w_cin + n) * w_cout + kk] : w[((size_t)kq * w_cin + kk) * w_cout + n];
            f[i] = v;
        }
        uint8_t *base = out + ((size_t)k * nchunks + c) * planes * slab;
        const size_t off = (size_t)n * 128 + (size_t)((g ^ (n & 7)) << 4);
        uint4 hi;
        hi.x = pack_bf16(f[0], f[1]);
        hi.y = pack_bf16(f[2], f[3]);
        hi.z = pack_bf16(f[4], f[5]);
        hi.w = pack_bf16(f[6], f[7]);
        *reinterpret_cast<uint4 *>(base + off) = hi;
        if (planes == 2) {
            float l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) l[i] = f[i] - __bfloat162float(__float2bfloat16_rn(f[i]));
            uint4 lo;
            lo.x = pack_bf16(l[0], l[1]);
            lo.y = pack_bf16(l[2], l[3]);
            lo.z = pack_bf16(l[4], l[5]);
            lo.w = pack_bf16(l[6], l[7]);
            *reinterpret_cast<uint4 *>(base + slab + off) = lo;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient on tcgen05:  dW[k][ci][co] += sum_j X[nbr[k, j]][ci] * dY[orow(j)][co]
//
// D[M = ci (128 lanes)][N = co] accumulates over K = rows j.  Both operands sit in shared memory as they
// sit in HBM — rows j, channels contiguous — which makes them MN-major operands: every 64-channel block is
// an [R rows][128 B] slab in the SWIZZLE_128B MN-major canonical layout (8-row groups 1024 B apart = SBO,
// 64-channel blocks one slab apart = LBO); one K = 16 step is two 8-row groups.
// One CTA = (offset k, 128-input-channel block, row split); it walks its rows in stages of 64, skipping
// 128-row tiles whose mask says no neighbour at k, and finally adds its partial dW tile with vector reds.
constexpr int WG_R = 64;                  // rows (GEMM-K) per stage
constexpr int WG_SLAB = WG_R * 128;       // bytes of one 64-channel block of one plane

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // next 64-element block along M/N
    d |= (uint64_t)(1024 >> 4) << 32;                  // next 8-row group along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct WgParams {
    const float *x;
    int ldx;
    const int32_t *nbr;
    int n_rows, kvol;
    const float *dy;
    int ldy;
    const int32_t *out_rows;
    float *dw;
    int cin, cout, mblks, splits, rows_per_split;
    const uint32_t *tile_mask;
    int npad, a_bytes, b_bytes, stage_bytes, stages, tmem_cols;
};

template <int PASSES>
__global__ void __launch_bounds__(THREADS, 1) k_wgrad_tc(WgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tmem;
    __shared__ uint32_t tmem_base_s;
    constexpr int NPL = PASSES == 3 ? 2 : 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);

    int b = blockIdx.x;
    const int split = b % p.splits;
    b /= p.splits;
    const int mblk = b % p.mblks;
    const int k = b / p.mblks;
    const int r_begin = split * p.rows_per_split;
    const int r_end = min(p.n_rows, r_begin + p.rows_per_split);
    const int ci0 = mblk * 128;
    const int m_valid = min(128, p.cin - ci0);

    // number of executed stages: identical in every role
    auto stage_active = [&](int r0) -> bool {
        return p.tile_mask == nullptr || ((p.tile_mask[r0 >> 7] >> k) & 1u);
    };
    int n_exec = 0;
    for (int r0 = r_begin; r0 < r_end; r0 += WG_R) n_exec += stage_active(r0) ? 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), PROD_WARPS);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        mbar_init(smem_u32(&bar_tmem), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < PROD_WARPS) {
        const int GB = p.cout / 8;  // 8-channel groups per dY row
        int item = 0;
        for (int r0 = r_begin; r0 < r_end; r0 += WG_R) {
            if (!stage_active(r0)) continue;
            const int s = item % p.stages;
            const uint32_t par = (item / p.stages) & 1;
            ++item;
            mbar_wait(smem_u32(&bar_empty[s]), par ^ 1, 10);
            uint8_t *a_hi = smem + (size_t)s * p.stage_bytes;
            uint8_t *a_lo = a_hi + p.a_bytes;
            uint8_t *b_hi = a_hi + NPL * p.a_bytes;
            uint8_t *b_lo = b_hi + p.b_bytes;
            // ---- A: 64 gathered rows x 128 input channels (16 groups of 8 per row)
            float4 va[4][2];
            int arow[4], agrp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int it = i * 256 + tid;
                arow[i] = it >> 4;
                agrp[i] = it & 15;
                const int j = r0 + arow[i];
                int src = -1;
                if (j < r_end) src = __ldg(p.nbr + (size_t)k * p.n_rows + j);
                const int c = ci0 + agrp[i] * 8;
                if (src >= 0 && c < p.cin) {
                    const float4 *ptr = reinterpret_cast<const float4 *>(p.x + (size_t)src * p.ldx + c);
                    va[i][0] = __ldg(ptr);
                    va[i][1] = __ldg(ptr + 1);
                } else {
                    va[i][0] = va[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            auto store_split = [&](uint8_t *hi_base, uint8_t *lo_base, int row, int grp, const float4 &q0, const float4 &q1) {
                const uint32_t off = (uint32_t)(grp >> 3) * WG_SLAB + (uint32_t)row * 128u + (uint32_t)(((grp & 7) ^ (row & 7)) << 4);
                const float f[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                uint4 hi;
                hi.x = pack_bf16(f[0], f[1]);
                hi.y = pack_bf16(f[2], f[3]);
                hi.z = pack_bf16(f[4], f[5]);
                hi.w = pack_bf16(f[6], f[7]);
                *reinterpret_cast<uint4 *>(hi_base + off) = hi;
                if (PASSES == 3) {
                    float l[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) l[e] = f[e] - __bfloat162float(__float2bfloat16_rn(f[e]));
                    uint4 lo;
                    lo.x = pack_bf16(l[0], l[1]);
                    lo.y = pack_bf16(l[2], l[3]);
                    lo.z = pack_bf16(l[4], l[5]);
                    lo.w = pack_bf16(l[6], l[7]);
                    *reinterpret_cast<uint4 *>(lo_base + off) = lo;
                }
            };
            // ---- B: 64 rows of dY x cout channels, 1024 (row, group) items per round
            const int nb_items = WG_R * GB;
            for (int base = 0; base < nb_items; base += 1024) {
                float4 vb[4][2];
                int brow[4], bgrp[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int it = base + i * 256 + tid;
                    brow[i] = it / GB;
                    bgrp[i] = it % GB;
                    vb[i][0] = vb[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (it < nb_items) {
                        const int j = r0 + brow[i];
                        if (j < r_end) {
                            const int orow = p.out_rows ? p.out_rows[j] : j;
                            const float4 *ptr = reinterpret_cast<const float4 *>(p.dy + (size_t)orow * p.ldy + bgrp[i] * 8);
                            vb[i][0] = __ldg(ptr);
                            vb[i][1] = __ldg(ptr + 1);
                        }
                    }
                }
                if (base == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) store_split(a_hi, a_lo, arow[i], agrp[i], va[i][0], va[i][1]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (base + i * 256 + tid < nb_items) store_split(b_hi, b_lo, brow[i], bgrp[i], vb[i][0], vb[i][1]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
        }
        // ---- epilogue: lanes = input channel, columns = output channel
        if (n_exec > 0) {
            const int quarter = warp & 3, half = warp >> 2;
            const int ci = quarter * 32 + lane;
            mbar_wait(smem_u32(&bar_tmem), 0, 11);
            tc_fence_after();
            float *drow = p.dw + ((size_t)k * p.cin + ci0 + ci) * p.cout;
            for (int col = half * 16; col < p.cout; col += 16 * (PROD_WARPS / 4)) {
                float acc[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col, acc);
                if (ci < m_valid) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) atomicAdd(drow + col + e, acc[e]);
                }
            }
        }
    } else if (warp == PROD_WARPS) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.npad) | (1u << 15) | (1u << 16);  // both operands MN-major
            const uint32_t lbo = WG_SLAB;
            for (int item = 0; item < n_exec; ++item) {
                const int s = item % p.stages;
                const uint32_t par = (item / p.stages) & 1;
                mbar_wait(smem_u32(&bar_full[s]), par, 12);
                tc_fence_after();
                const uint32_t a_hi = smem_base + (uint32_t)s * p.stage_bytes;
                const uint32_t a_lo = a_hi + p.a_bytes;
                const uint32_t b_hi = a_hi + NPL * p.a_bytes;
                const uint32_t b_lo = b_hi + p.b_bytes;
#pragma unroll
                for (int kk = 0; kk < WG_R / 16; ++kk) {
                    const uint32_t adv = (uint32_t)kk * 2048u;  // two 8-row groups
                    const uint64_t da_hi = make_desc_mn_sw128(a_hi + adv, lbo), db_hi = make_desc_mn_sw128(b_hi + adv, lbo);
                    umma(tmem_base, da_hi, db_hi, idesc, (item | kk) != 0);
                    if (PASSES == 3) {
                        umma(tmem_base, make_desc_mn_sw128(a_lo + adv, lbo), db_hi, idesc, 1);
                        umma(tmem_base, da_hi, make_desc_mn_sw128(b_lo + adv, lbo), idesc, 1);
                    }
                }
                umma_commit(smem_u32(&bar_empty[s]));
            }
            if (n_exec > 0) umma_commit(smem_u32(&bar_tmem));
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

}  // namespace tc
}  // namespace us3d

using namespace us3d;

extern "C" {

long long us3d_spconv_packed_bytes(int kvol, int kdim, int ndim, int passes) {
    int nchunks = (kdim + tc::KC - 1) / tc::KC;
    return (long long)kvol * nchunks * (passes == 3 ? 2 : 1) * ndim * 128;
}

int us3d_spconv_tc_supported(int cin, int cout) {
    return cin >= 16 && cin % 16 == 0 && cout >= 16 && cout % 16 == 0 && cout <= 256;
}

int us3d_spconv_pack_weights(const float *w, int kvol, int cin, int cout, int transpose, int flip_k, int passes,
                             void *out, void *stream_) {
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "pack_weights: kvol %d out of range", kvol);
    US3D_CHECK_ARG(passes == 1 || passes == 3, "pack_weights: passes must be 1 or 3");
    const int kdim = transpose ? cout : cin, ndim = transpose ? cin : cout;
    US3D_CHECK_ARG(us3d_spconv_tc_supported(kdim, ndim), "pack_weights: unsupported channel counts %d -> %d", kdim, ndim);
    const int nchunks = (kdim + tc::KC - 1) / tc::KC;
    long long total = (long long)kvol * nchunks * ndim * 8;
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    tc::k_pack_weights<<<grid, 256, 0, (cudaStream_t)stream_>>>(w, kvol, kdim, ndim, cin, cout, transpose, flip_k,
                                                               passes == 3 ? 2 : 1, (uint8_t *)out);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_spconv_wgrad_tc_supported(int cin, int cout) { return cin >= 8 && cin % 8 == 0 && cout >= 16 && cout % 16 == 0 && cout <= 256; }

int us3d_spconv_wgrad_tc(const float *x, int ldx, const int32_t *nbr, int n_rows, int kvol, const float *dy, int ldy,
                         const int32_t *out_rows, float *dw, int cin, int cout, int passes, const uint32_t *tile_mask,
                         void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "spconv_wgrad_tc: kvol %d out of range", kvol);
    US3D_CHECK_ARG(passes == 1 || passes == 3, "spconv_wgrad_tc: passes must be 1 or 3");
    US3D_CHECK_ARG(us3d_spconv_wgrad_tc_supported(cin, cout), "spconv_wgrad_tc: unsupported channel counts %d -> %d", cin, cout);
    US3D_CHECK_ARG(ldx % 4 == 0 && ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0,
                   "spconv_wgrad_tc: x and dy must be 16-byte aligned with leading dimensions %% 4 == 0");
    if (n_rows == 0) return 0;
    tc::WgParams p;
    p.x = x; p.ldx = ldx; p.nbr = nbr; p.n_rows = n_rows; p.kvol = kvol; p.dy = dy; p.ldy = ldy; p.out_rows = out_rows;
    p.dw = dw; p.cin = cin; p.cout = cout; p.tile_mask = tile_mask;
    p.mblks = ceil_div(cin, 128);
    int splits = (2 * num_sms()) / (kvol * p.mblks);  // whole waves of one CTA per SM: never spill into a third wave
    int max_splits = ceil_div(n_rows, 1024);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.rows_per_split = ceil_div(ceil_div(n_rows, splits), 128) * 128;
    p.splits = ceil_div(n_rows, p.rows_per_split);
    p.npad = ceil_div(cout, 64) * 64;
    const int npl = passes == 3 ? 2 : 1;
    p.a_bytes = 2 * tc::WG_SLAB;
    p.b_bytes = (p.npad / 64) * tc::WG_SLAB;
    p.stage_bytes = npl * (p.a_bytes + p.b_bytes);
    int stages = (200 * 1024) / p.stage_bytes;
    if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
    US3D_CHECK_ARG(stages >= 2, "spconv_wgrad_tc: stage of %d bytes does not fit twice in shared memory", p.stage_bytes);
    p.stages = stages;
    p.tmem_cols = p.npad < 32 ? 32 : p.npad;
    const size_t smem = (size_t)stages * p.stage_bytes + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        US3D_CUDA(cudaFuncSetAttribute(tc::k_wgrad_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        US3D_CUDA(cudaFuncSetAttribute(tc::k_wgrad_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        attr_done = true;
    }
    const int grid = kvol * p.mblks * p.splits;
    if (passes == 3)
        tc::k_wgrad_tc<3><<<grid, tc::THREADS, smem, st>>>(p);
    else
        tc::k_wgrad_tc<1><<<grid, tc::THREADS, smem, st>>>(p);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_spconv_gather_tc(const float *x, int ldx, const int32_t *nbr, int n_rows, int kvol, const void *wpack, int cin,
                          int cout, int passes, const float *bias, const int32_t *out_rows, float *y, int ldy,
                          int accumulate, const uint32_t *tile_mask, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "spconv_gather_tc: kvol %d out of range", kvol);
    US3D_CHECK_ARG(passes == 1 || passes == 3, "spconv_gather_tc: passes must be 1 or 3");
    US3D_CHECK_ARG(us3d_spconv_tc_supported(cin, cout), "spconv_gather_tc: unsupported channel counts %d -> %d", cin, cout);
    US3D_CHECK_ARG(ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "spconv_gather_tc: x must be 16-byte aligned with ldx %% 4 == 0");
    US3D_CHECK_ARG(ldy >= cout && ldx >= cin, "spconv_gather_tc: bad leading dimensions");
    if (n_rows == 0) return 0;
    tc::Params p;
    p.x = x; p.ldx = ldx; p.nbr = nbr; p.n_rows = n_rows; p.kvol = kvol;
    p.wpack = (const uint8_t *)wpack; p.cin = cin; p.cout = cout; p.nchunks = (cin + tc::KC - 1) / tc::KC;
    p.bias = bias; p.out_rows = out_rows; p.y = y; p.ldy = ldy; p.accumulate = accumulate; p.tile_mask = tile_mask;
    const int npl = passes == 3 ? 2 : 1;
    p.b_bytes = cout * 128;
    p.stage_bytes = npl * (tc::A_BYTES + p.b_bytes);
    int stages = (200 * 1024) / p.stage_bytes;
    if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
    US3D_CHECK_ARG(stages >= 2, "spconv_gather_tc: stage of %d bytes does not fit twice in shared memory", p.stage_bytes);
    // keep two CTAs per SM when that still leaves >= 2 stages each
    if (stages > 2 && 2 * (2 * p.stage_bytes + 2048) <= 220 * 1024) {
        int s2 = (108 * 1024) / p.stage_bytes;
        stages = s2 < 2 ? 2 : (s2 > tc::MAX_STAGES ? tc::MAX_STAGES : s2);
    }
    p.stages = stages;
    int cols = 32;
    while (cols < cout) cols <<= 1;
    p.tmem_cols = cols;
    const size_t smem = (size_t)stages * p.stage_bytes + 1024;
    const int tiles = ceil_div(n_rows, tc::M);
    static bool attr_done = false;
    if (!attr_done) {
        US3D_CUDA(cudaFuncSetAttribute(tc::k_spconv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        US3D_CUDA(cudaFuncSetAttribute(tc::k_spconv_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        attr_done = true;
    }
    if (passes == 3)
        tc::k_spconv_tc<3><<<tiles, tc::THREADS, smem, st>>>(p);
    else
        tc::k_spconv_tc<1><<<tiles, tc::THREADS, smem, st>>>(p);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
