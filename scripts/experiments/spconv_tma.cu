// Sparse convolution, Blackwell-native data path: TMA row gather -> tcgen05.mma -> TMEM -> epilogue.
//
//     Y[orow(j)] = sum_k X[nbr[k, j]] · W'[k]          persistent CTAs, one 128-row output tile at a time
//
// Activations are consumed as bf16 planes (hi, and lo = x - hi for the fp32-faithful three-term split)
// produced by us3d_split_bf16.  For every (tile, active offset k, 64-channel chunk) stage:
//
//   warp 0  (producer, all 32 lanes)   lane l issues ONE cp.async.bulk.tensor...tile::gather4 per plane: the
//           four neighbour rows 4l..4l+3 of the tile, 64 channels (128 B) each, land in rows 4l..4l+3 of the
//           128 x 128 B stage buffer already in the K-major SWIZZLE_128B layout tcgen05 expects; absent
//           neighbours are passed as row -1, which the TMA unit zero-fills without touching memory.  Lane 0
//           also streams the pre-packed weight slab with cp.async.bulk.  All of it completes on one mbarrier.
//   warp 1  (MMA, one lane)            tcgen05.mma M=128, N=Cout, K=16, kind::f16 into one of two TMEM
//           accumulators; tcgen05.commit frees the stage / publishes the accumulator.
//   warps 2-5 (epilogue)               tcgen05.ld -> (+bias, +=) -> fp32 rows; runs while the next tile's main
//           loop fills the other accumulator.
//
// No register staging, no generic-proxy writes to the operand buffers, no atomics.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace us3d {
namespace tma {

constexpr int M = 128;
constexpr int KC = 64;
constexpr int THREADS = 192;
constexpr int MAX_STAGES = 6;
constexpr int A_BYTES = M * 128;

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
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("us3d spconv_tma: mbarrier wait timed out (role %d, block %d, thread %d, parity %u)\n", who, blockIdx.x,
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
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap *map, int col, int row, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(col), "r"(row), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {  // K-major, SWIZZLE_128B
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {  // MN-major, SWIZZLE_128B
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {  // bf16 x bf16 -> fp32, M = 128
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

// ------------------------------------------------------------------------------------------------- forward / dgrad
struct Params {
    const int32_t *nbr;
    int n_rows, kvol, n_tiles;
    const uint8_t *wpack;
    int cin, cout, nchunks;
    const float *bias;
    const int32_t *out_rows;
    float *y;
    int ldy, accumulate;
    const uint32_t *tile_mask;
    int stages, b_bytes, stage_bytes, acc_cols;
};

template <int PASSES>
__global__ void __launch_bounds__(THREADS, 1)
k_spconv_tma(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    constexpr int NPL = PASSES == 3 ? 2 : 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t all_k = p.kvol >= 32 ? 0xFFFFFFFFu : ((1u << p.kvol) - 1u);

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);   // the producer's expect_tx arrive; TMA completes the bytes
            mbar_init(smem_u32(&bar_empty[s]), 1);  // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bar_acc_full[b]), 1);   // tcgen05.commit after the tile's last MMA
            mbar_init(smem_u32(&bar_acc_empty[b]), 4);  // the four epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)(2 * p.acc_cols))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        const uint32_t stage_tx = (uint32_t)(NPL * (A_BYTES + p.b_bytes));
        int item = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const uint32_t kmask = p.tile_mask ? (p.tile_mask[tile] & all_k) : all_k;
            const int row = tile * M + 4 * lane;
            for (int k = 0; k < p.kvol; ++k) {
                if (!((kmask >> k) & 1u)) continue;
                const int32_t *nk = p.nbr + (size_t)k * p.n_rows;
                const int i0 = row + 0 < p.n_rows ? __ldg(nk + row + 0) : -1;
                const int i1 = row + 1 < p.n_rows ? __ldg(nk + row + 1) : -1;
                const int i2 = row + 2 < p.n_rows ? __ldg(nk + row + 2) : -1;
                const int i3 = row + 3 < p.n_rows ? __ldg(nk + row + 3) : -1;
                for (int c = 0; c < p.nchunks; ++c, ++item) {
                    const int s = item % p.stages;
                    const uint32_t par = (item / p.stages) & 1;
                    mbar_wait(smem_u32(&bar_empty[s]), par ^ 1, 0);
                    const uint32_t bar = smem_u32(&bar_full[s]);
                    const uint32_t a_hi = smem_base + (uint32_t)s * p.stage_bytes;
                    if (lane == 0) {
                        mbar_arrive_expect_tx(bar, stage_tx);
                        const uint32_t b_dst = a_hi + NPL * A_BYTES;
                        const uint8_t *src = p.wpack + ((size_t)k * p.nchunks + c) * NPL * (size_t)p.b_bytes;
                        bulk_g2s(b_dst, src, (uint32_t)p.b_bytes, bar);
                        if (PASSES == 3) bulk_g2s(b_dst + p.b_bytes, src + p.b_bytes, (uint32_t)p.b_bytes, bar);
                    }
                    __syncwarp();
                    const uint32_t dst = a_hi + (uint32_t)(4 * lane) * 128u;
                    tma_gather4(dst, &map_hi, c * KC, i0, i1, i2, i3, bar);
                    if (PASSES == 3) tma_gather4(dst + A_BYTES, &map_lo, c * KC, i0, i1, i2, i3, bar);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.cout);
            int item = 0, titer = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++titer) {
                const uint32_t kmask = p.tile_mask ? (p.tile_mask[tile] & all_k) : all_k;
                const int n_items = __popc(kmask) * p.nchunks;
                const int buf = titer & 1;
                mbar_wait(smem_u32(&bar_acc_empty[buf]), ((titer >> 1) & 1) ^ 1, 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * p.acc_cols);
                for (int i = 0; i < n_items; ++i, ++item) {
                    const int s = item % p.stages;
                    const uint32_t par = (item / p.stages) & 1;
                    const int c = i % p.nchunks;
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
                        const uint64_t adv = (uint64_t)(kk * 2);
                        umma(acc, da_hi + adv, db_hi + adv, idesc, (i | kk) != 0);
                        if (PASSES == 3) {
                            umma(acc, da_lo + adv, db_hi + adv, idesc, 1);
                            umma(acc, da_hi + adv, db_lo + adv, idesc, 1);
                        }
                    }
                    umma_commit(smem_u32(&bar_empty[s]));
                }
                if (n_items > 0)
                    umma_commit(smem_u32(&bar_acc_full[buf]));
                else
                    mbar_arrive(smem_u32(&bar_acc_full[buf]));
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int quarter = warp & 3;  // TMEM lane quarter this warp may read
        const bool vec = (p.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        int titer = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++titer) {
            const uint32_t kmask = p.tile_mask ? (p.tile_mask[tile] & all_k) : all_k;
            const bool has_acc = kmask != 0;
            const int buf = titer & 1;
            const int j = tile * M + quarter * 32 + lane;
            const bool row_ok = j < p.n_rows;
            float *yrow = nullptr;
            if (row_ok) yrow = p.y + (size_t)(p.out_rows ? p.out_rows[j] : j) * p.ldy;
            mbar_wait(smem_u32(&bar_acc_full[buf]), (titer >> 1) & 1, 3);
            tc_fence_after();
            const uint32_t acc_addr = tmem_base + (uint32_t)(buf * p.acc_cols) + ((uint32_t)(quarter * 32) << 16);
            for (int col = 0; col < p.cout; col += 16) {
                float acc[16];
                if (has_acc) {
                    tmem_ld16(acc_addr + (uint32_t)col, acc);
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[buf]));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * p.acc_cols)) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------- cp.async producers
// Same pipeline, but the 128 gathered rows of a stage are fetched with 16-byte cp.async (LDGSTS) by four
// producer warps instead of TMA gather4.  Measured on B200: one tile::gather4 costs ~80 SM cycles in the TMA
// unit (~6 B/clk/SM), an LDGSTS warp instruction moves 512 B in ~8 cycles — the gather is an order of magnitude
// faster through the LSU.  Completion: each producer thread commits one cp.async group per stage, waits until the
// group of LAG stages ago has landed (cp.async.wait_group), makes it visible to the tensor-core proxy
// (fence.proxy.async) and its warp leader arrives on that stage's mbarrier.
constexpr int CP_PROD_WARPS = 4;
constexpr int CP_THREADS = (CP_PROD_WARPS + 1 + 4) * 32;  // producers, MMA warp, four epilogue warps

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct CpParams {
    const __nv_bfloat16 *x_hi, *x_lo;  // [n_in, cin] planes
    const int32_t *nbr;
    int n_rows, kvol, n_tiles;
    const uint8_t *wpack;
    int cin, cout, nchunks;
    const float *bias;
    const int32_t *out_rows;
    float *y;
    int ldy, accumulate;
    const uint32_t *tile_mask;
    int stages, b_bytes, stage_bytes, acc_cols;
};

template <int PASSES, int LAG>
__global__ void __launch_bounds__(CP_THREADS, 1) k_spconv_cp(CpParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    constexpr int NPL = PASSES == 3 ? 2 : 1;
    constexpr int MMA_WARP = CP_PROD_WARPS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t all_k = p.kvol >= 32 ? 0xFFFFFFFFu : ((1u << p.kvol) - 1u);

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), CP_PROD_WARPS + 1);  // producer warps + the weight slab's expect_tx arrive
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bar_acc_full[b]), 1);
            mbar_init(smem_u32(&bar_acc_empty[b]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)(2 * p.acc_cols))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < CP_PROD_WARPS) {
        // ------------------------------------------------------------------ gather producers (128 threads)
        constexpr int RPT = M * 8 / (CP_PROD_WARPS * 32);  // rows per thread: 8
        const int grp = tid & 7;      // 16-byte chunk (8 channels) inside the 64-channel stage
        const int rbase = tid >> 3;   // rows rbase + 16 i
        int item = 0;                 // stages issued so far (ring position)
        int signalled = 0;            // stages whose full-barrier arrive has been done
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const uint32_t kmask = p.tile_mask ? (p.tile_mask[tile] & all_k) : all_k;
            const int tile0 = tile * M;
            for (int k = 0; k < p.kvol; ++k) {
                if (!((kmask >> k) & 1u)) continue;
                int idx[RPT];
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const int j = tile0 + rbase + 16 * i;
                    idx[i] = j < p.n_rows ? __ldg(p.nbr + (size_t)k * p.n_rows + j) : -1;
                }
                for (int c = 0; c < p.nchunks; ++c, ++item) {
                    const int s = item % p.stages;
                    const uint32_t par = (item / p.stages) & 1;
                    mbar_wait(smem_u32(&bar_empty[s]), par ^ 1, 0);
                    const uint32_t a_hi = smem_base + (uint32_t)s * p.stage_bytes;
                    if (tid == 0) {
                        const uint32_t bar = smem_u32(&bar_full[s]);
                        mbar_arrive_expect_tx(bar, (uint32_t)(NPL * p.b_bytes));
                        const uint32_t b_dst = a_hi + NPL * A_BYTES;
                        const uint8_t *src = p.wpack + ((size_t)k * p.nchunks + c) * NPL * (size_t)p.b_bytes;
                        bulk_g2s(b_dst, src, (uint32_t)p.b_bytes, bar);
                        if (PASSES == 3) bulk_g2s(b_dst + p.b_bytes, src + p.b_bytes, (uint32_t)p.b_bytes, bar);
                    }
                    const int c0 = c * KC + grp * 8;
                    const bool col_ok = c0 < p.cin;  // chunk columns past cin are zero-filled
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const int r = rbase + 16 * i;
                        const uint32_t dst = a_hi + (uint32_t)r * 128u + (uint32_t)((grp ^ (r & 7)) << 4);
                        const bool ok = idx[i] >= 0 && col_ok;
                        const size_t off = ok ? (size_t)idx[i] * p.cin + c0 : 0;
                        cp_async16(dst, p.x_hi + off, ok ? 16u : 0u);
                        if (PASSES == 3) cp_async16(dst + A_BYTES, p.x_lo + off, ok ? 16u : 0u);
                    }
                    cp_async_commit();
                    if (item - signalled >= LAG) {
                        cp_async_wait<LAG>();
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bar_full[signalled % p.stages]));
                        ++signalled;
                    }
                }
            }
        }
        // drain
        cp_async_wait<0>();
        fence_proxy_async();
        __syncwarp();
        for (; signalled < item; ++signalled)
            if (lane == 0) mbar_arrive(smem_u32(&bar_full[signalled % p.stages]));
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.cout);
            int item = 0, titer = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++titer) {
                const uint32_t kmask = p.tile_mask ? (p.tile_mask[tile] & all_k) : all_k;
                const int n_items = __popc(kmask) * p.nchunks;
                const int buf = titer & 1;
                mbar_wait(smem_u32(&bar_acc_empty[buf]), ((titer >> 1) & 1) ^ 1, 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * p.acc_cols);
                for (int i = 0; i < n_items; ++i, ++item) {
                    const int s = item % p.stages;
                    const uint32_t par = (item / p.stages) & 1;
                    const int c = i % p.nchunks;
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
                        const uint64_t adv = (uint64_t)(kk * 2);
                        umma(acc, da_hi + adv, db_hi + adv, idesc, (i | kk) != 0);
                        if (PASSES == 3) {
                            umma(acc, da_lo + adv, db_hi + adv, idesc, 1);
                            umma(acc, da_hi + adv, db_lo + adv, idesc, 1);
                        }
                    }
                    umma_commit(smem_u32(&bar_empty[s]));
                }
                if (n_items > 0)
                    umma_commit(smem_u32(&bar_acc_full[buf]));
                else
                    mbar_arrive(smem_u32(&bar_acc_full[buf]));
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 5..8)
        const int quarter = warp & 3;
        const bool vec = (p.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        int titer = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++titer) {
            const uint32_t kmask = p.tile_mask ? (p.tile_mask[tile] & all_k) : all_k;
            const bool has_acc = kmask != 0;
            const int buf = titer & 1;
            const int j = tile * M + quarter * 32 + lane;
            const bool row_ok = j < p.n_rows;
            float *yrow = nullptr;
            if (row_ok) yrow = p.y + (size_t)(p.out_rows ? p.out_rows[j] : j) * p.ldy;
            mbar_wait(smem_u32(&bar_acc_full[buf]), (titer >> 1) & 1, 3);
            tc_fence_after();
            const uint32_t acc_addr = tmem_base + (uint32_t)(buf * p.acc_cols) + ((uint32_t)(quarter * 32) << 16);
            for (int col = 0; col < p.cout; col += 16) {
                float acc[16];
                if (has_acc) {
                    tmem_ld16(acc_addr + (uint32_t)col, acc);
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[buf]));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * p.acc_cols)) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------- helpers
// fp32 rows -> bf16 hi plane (+ lo plane = x - hi), row-major [n, c], c % 8 == 0
__global__ void __launch_bounds__(256) k_split_bf16(const float *__restrict__ x, int ldx, int n, int c, uint4 *__restrict__ hi,
                                                    uint4 *__restrict__ lo) {
    const int g = c / 8;
    const long long total = (long long)n * g;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / g), q = (int)(e % g);
        const float4 *src = reinterpret_cast<const float4 *>(x + (size_t)r * ldx + q * 8);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint4 h;
        h.x = pack_bf16(f[0], f[1]);
        h.y = pack_bf16(f[2], f[3]);
        h.z = pack_bf16(f[4], f[5]);
        h.w = pack_bf16(f[6], f[7]);
        hi[e] = h;
        if (lo != nullptr) {
            float l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) l[i] = f[i] - __bfloat162float(__float2bfloat16_rn(f[i]));
            uint4 w;
            w.x = pack_bf16(l[0], l[1]);
            w.y = pack_bf16(l[2], l[3]);
            w.z = pack_bf16(l[4], l[5]);
            w.w = pack_bf16(l[6], l[7]);
            lo[e] = w;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void *ptr = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 2-D bf16 row-major [rows, cols] map with a (64 x box_rows) box and the 128-byte swizzle
static int make_map(CUtensorMap *map, const void *base, int rows, int cols, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return -4;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with %d for a [%d, %d] bf16 plane", (int)rc, rows, cols);
        return -4;
    }
    return 0;
}

}  // namespace tma
}  // namespace us3d

using namespace us3d;

extern "C" {

int us3d_split_bf16(const float *x, int ldx, int n, int c, void *hi, void *lo, void *stream_) {
    US3D_CHECK_ARG(n >= 0 && c > 0 && c % 8 == 0 && ldx % 4 == 0 && ldx >= c, "split_bf16: need c %% 8 == 0 and ldx %% 4 == 0");
    US3D_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "split_bf16: x must be 16-byte aligned");
    if (n == 0) return 0;
    long long total = (long long)n * (c / 8);
    long long grid = (total + 255) / 256;
    if (grid > (long long)num_sms() * 16) grid = (long long)num_sms() * 16;
    tma::k_split_bf16<<<(int)grid, 256, 0, (cudaStream_t)stream_>>>(x, ldx, n, c, (uint4 *)hi, (uint4 *)lo);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

template <int PASSES>
static void launch_cp(int lag, int grid, size_t smem, cudaStream_t st, const tma::CpParams &p) {
    if (lag >= 3)
        tma::k_spconv_cp<PASSES, 3><<<grid, tma::CP_THREADS, smem, st>>>(p);
    else if (lag == 2)
        tma::k_spconv_cp<PASSES, 2><<<grid, tma::CP_THREADS, smem, st>>>(p);
    else
        tma::k_spconv_cp<PASSES, 1><<<grid, tma::CP_THREADS, smem, st>>>(p);
}

extern "C" {

int us3d_spconv_gather_cp(const void *x_hi, const void *x_lo, int n_in, const int32_t *nbr, int n_rows, int kvol,
                          const void *wpack, int cin, int cout, int passes, const float *bias, const int32_t *out_rows,
                          float *y, int ldy, int accumulate, const uint32_t *tile_mask, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "spconv_gather_cp: kvol %d out of range", kvol);
    US3D_CHECK_ARG(passes == 1 || passes == 3, "spconv_gather_cp: passes must be 1 or 3");
    US3D_CHECK_ARG(us3d_spconv_tc_supported(cin, cout), "spconv_gather_cp: unsupported channel counts %d -> %d", cin, cout);
    US3D_CHECK_ARG(x_hi != nullptr && (passes == 1 || x_lo != nullptr), "spconv_gather_cp: missing activation plane");
    US3D_CHECK_ARG((reinterpret_cast<uintptr_t>(x_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_lo) & 15) == 0,
                   "spconv_gather_cp: planes must be 16-byte aligned");
    US3D_CHECK_ARG(ldy >= cout && n_in > 0, "spconv_gather_cp: bad sizes");
    if (n_rows == 0) return 0;
    tma::CpParams p;
    p.x_hi = (const __nv_bfloat16 *)x_hi; p.x_lo = (const __nv_bfloat16 *)x_lo;
    p.nbr = nbr; p.n_rows = n_rows; p.kvol = kvol; p.n_tiles = ceil_div(n_rows, tma::M);
    p.wpack = (const uint8_t *)wpack; p.cin = cin; p.cout = cout; p.nchunks = ceil_div(cin, tma::KC);
    p.bias = bias; p.out_rows = out_rows; p.y = y; p.ldy = ldy; p.accumulate = accumulate; p.tile_mask = tile_mask;
    const int npl = passes == 3 ? 2 : 1;
    p.b_bytes = cout * 128;
    p.stage_bytes = npl * (tma::A_BYTES + p.b_bytes);
    int stages = (208 * 1024) / p.stage_bytes;
    if (stages > tma::MAX_STAGES) stages = tma::MAX_STAGES;
    US3D_CHECK_ARG(stages >= 2, "spconv_gather_cp: stage of %d bytes does not fit twice in shared memory", p.stage_bytes);
    p.stages = stages;
    int cols = 32;
    while (cols < cout) cols <<= 1;
    p.acc_cols = cols;
    const size_t smem = (size_t)stages * p.stage_bytes + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        US3D_CUDA(cudaFuncSetAttribute(tma::k_spconv_cp<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        US3D_CUDA(cudaFuncSetAttribute(tma::k_spconv_cp<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        US3D_CUDA(cudaFuncSetAttribute(tma::k_spconv_cp<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        US3D_CUDA(cudaFuncSetAttribute(tma::k_spconv_cp<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        US3D_CUDA(cudaFuncSetAttribute(tma::k_spconv_cp<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        US3D_CUDA(cudaFuncSetAttribute(tma::k_spconv_cp<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        attr_done = true;
    }
    const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
    const int lag = stages - 1;
    if (passes == 3)
        launch_cp<3>(lag, grid, smem, st, p);
    else
        launch_cp<1>(lag, grid, smem, st, p);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_spconv_gather_tma(const void *x_hi, const void *x_lo, int n_in, const int32_t *nbr, int n_rows, int kvol,
                           const void *wpack, int cin, int cout, int passes, const float *bias, const int32_t *out_rows,
                           float *y, int ldy, int accumulate, const uint32_t *tile_mask, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "spconv_gather_tma: kvol %d out of range", kvol);
    US3D_CHECK_ARG(passes == 1 || passes == 3, "spconv_gather_tma: passes must be 1 or 3");
    US3D_CHECK_ARG(us3d_spconv_tc_supported(cin, cout), "spconv_gather_tma: unsupported channel counts %d -> %d", cin, cout);
    US3D_CHECK_ARG(x_hi != nullptr && (passes == 1 || x_lo != nullptr), "spconv_gather_tma: missing activation plane");
    US3D_CHECK_ARG(ldy >= cout && n_in > 0, "spconv_gather_tma: bad sizes");
    if (n_rows == 0) return 0;
    CUtensorMap map_hi, map_lo;
    int rc = tma::make_map(&map_hi, x_hi, n_in, cin, 1);
    if (rc) return rc;
    rc = tma::make_map(&map_lo, passes == 3 ? x_lo : x_hi, n_in, cin, 1);
    if (rc) return rc;
    tma::Params p;
    p.nbr = nbr; p.n_rows = n_rows; p.kvol = kvol; p.n_tiles = ceil_div(n_rows, tma::M);
    p.wpack = (const uint8_t *)wpack; p.cin = cin; p.cout = cout; p.nchunks = ceil_div(cin, tma::KC);
    p.bias = bias; p.out_rows = out_rows; p.y = y; p.ldy = ldy; p.accumulate = accumulate; p.tile_mask = tile_mask;
    const int npl = passes == 3 ? 2 : 1;
    p.b_bytes = cout * 128;
    p.stage_bytes = npl * (tma::A_BYTES + p.b_bytes);
    int stages = (208 * 1024) / p.stage_bytes;
    if (stages > tma::MAX_STAGES) stages = tma::MAX_STAGES;
    US3D_CHECK_ARG(stages >= 2, "spconv_gather_tma: stage of %d bytes does not fit twice in shared memory", p.stage_bytes);
    p.stages = stages;
    int cols = 32;
    while (cols < cout) cols <<= 1;
    p.acc_cols = cols;  // two accumulators: 2 * cols <= 512
    const size_t smem = (size_t)stages * p.stage_bytes + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        US3D_CUDA(cudaFuncSetAttribute(tma::k_spconv_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        US3D_CUDA(cudaFuncSetAttribute(tma::k_spconv_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        attr_done = true;
    }
    int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
    if (passes == 3)
        tma::k_spconv_tma<3><<<grid, tma::THREADS, smem, st>>>(map_hi, map_lo, p);
    else
        tma::k_spconv_tma<1><<<grid, tma::THREADS, smem, st>>>(map_hi, map_lo, p);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
