// Sparse convolution forward / input-gradient — the production tcgen05 kernel ("multi-tile").
//
//     Y[orow(j)] = sum_k X[nbr[k, j]] · W'[k]
//
// Every tcgen05 variant measured on B200 (register-staged, cp.async, TMA gather4) ran into the same wall: the
// L2 -> SM fabric tops out near 6-7 TB/s, and with one 128-row tile per CTA the weight slabs re-streamed by
// every tile are 60 % of that traffic.  This kernel therefore keeps T = 512 / acc_cols (up to 4) output tiles
// in flight per CTA: one weight slab (offset k, 64-channel chunk) is loaded ONCE into a B ring slot and
// multiplied against the gathered A tiles of all T tiles, each accumulating into its own TMEM accumulator.
//
//   warps 0-7    A producers: 16-byte cp.async (LDGSTS, PTX ignore-src predicate for absent neighbours, one 32 x 32 -> 64 bit
//                multiply-add per row address) from the bf16 planes (hi, + lo for the three-term split) straight into K-major
//                SWIZZLE_128B slots; completion by cp.async.mbarrier.arrive.noinc
//   warp 8       B loader: cp.async.bulk of the pre-packed, pre-swizzled weight slab
//   warps 9,15-17 MMA issuers, one per tile of the group: tcgen05.mma M=128, N=Cout (2 Cout for the fused operand), K=16;
//                tcgen05.commit frees slots
//   warps 10-13, 18-21  two epilogue sets (set e: tiles e, e + 2): tcgen05.ld -> (+bias, +=) -> fp32 rows, BatchNorm column sums
//   warp 14      neighbour-index ring
//
// Round 2:
//   * warp 10 streams the neighbour indices of the next NIDX (super-tile, offset) pairs into a shared-memory ring with
//     4-byte cp.async (mbarrier completion): the producers no longer expose one L2 round trip per offset, and no index
//     lives in a register across an item (a register double buffer was measured 27 % SLOWER: ptxas spilled the prefetched
//     values, which turns every prefetch into a blocking load);
//   * LAG == 0: completion by cp.async.mbarrier.arrive.noinc — a producer never waits for its own copies, the whole A
//     ring can be in flight (the MMA warp crosses generic -> async proxy with fence.proxy.async after its wait);
//   * FUSE (three-term mode, Cout <= 128): the weight slab [W_hi rows | W_lo rows] is one K-major B operand of
//     N = 2 Cout rows, so X_hi W_hi and X_hi W_lo are ONE tcgen05.mma (A read from shared memory once, 10 KB of operands
//     per 96 issue cycles instead of 2 x 7 KB per 2 x 48) landing in column groups [0, Cout) and [Cout, 2 Cout) that the
//     epilogue adds; X_lo W_hi follows with N = Cout.
//
// Replaces MinkowskiConvolution / MinkowskiConvolutionTranspose forward and input gradient
// (/root/reference/models/modules/common.py:146-155, 179-188).
#include "common.cuh"
#include "tc_common.cuh"

namespace us3d {
namespace mt {

using namespace tcx;

constexpr int M = 128;
constexpr int KC = 64;
constexpr int A_PLANE = M * 128;  // bytes of one plane of one A slot
#ifndef US3D_MT_PROD_WARPS
#define US3D_MT_PROD_WARPS 8
#endif
constexpr int PROD_WARPS = US3D_MT_PROD_WARPS;
constexpr int B_WARP = PROD_WARPS, MMA_WARP = PROD_WARPS + 1, EPI_WARP0 = PROD_WARPS + 2;
constexpr int IDX_WARP = PROD_WARPS + 6;
constexpr int MMA_WARP_EXTRA0 = PROD_WARPS + 7;  // MMA warps 1 .. MAX_T-1 (one MMA-issuing warp per tile of the group)
constexpr int EPI2_WARP0 = PROD_WARPS + 10;      // second epilogue set (PROD_WARPS % 4 == 0: warp % 4 covers the four TMEM lane quarters)
constexpr int EPI_SETS = 2;                      // epilogue set e reads out tiles t = e, e + 2 of a group
constexpr int THREADS = (PROD_WARPS + 2 + 4 + 1 + 3 + 4) * 32;
static_assert(PROD_WARPS % 4 == 0, "epilogue warps are placed by warp % 4");
constexpr int MAX_A = 8, MAX_B = 3, MAX_T = 4;
constexpr int NIDX = 4;  // stages of the neighbour-index ring (one stage = the indices of one offset for the T tiles)

// BatchNorm statistics of the output, folded into the epilogue (ws == NULL: off).  ws = 2 cout doubles (column sums, sums of
// squares) + a ticket, zero on entry and zero again on exit — the same scratch and the same finalisation as us3d_bn_stats_fused.
struct BnFuse {
    double *ws;
    float *mean, *invstd, *running_mean, *running_var;
    long long *num_batches_tracked;
    float eps, momentum;
};

// Column sums over the 32 rows of a warp: lane l holds v[0..15] of its row; on return v[0] of lane l is the sum of column
// ((l >> 1) & 15 read as bits 4..1 of l: 8 (l>>4 & 1) + 4 (l>>3 & 1) + 2 (l>>2 & 1) + (l>>1 & 1)) over all 32 lanes.
// A halving butterfly: 8 + 4 + 2 + 1 + 1 shuffles.
__device__ __forceinline__ void warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
    for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const bool up = (lane & bit) != 0;
            const float keep = up ? v[i + w] : v[i];
            const float send = up ? v[i] : v[i + w];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int warp_colsum16_col(int lane) {
    return 8 * ((lane >> 4) & 1) + 4 * ((lane >> 3) & 1) + 2 * ((lane >> 2) & 1) + ((lane >> 1) & 1);
}

// The CTA's column sums (shared memory, [2][c] floats) go to the global fp64 scratch; the CTA that takes the last ticket
// finalises mean / invstd / running statistics and zeroes the scratch (fused::k_bn_stats_fused does the same after its own sums).
__device__ __forceinline__ void bn_publish_and_finalize(const BnFuse &bn, const float *s_sum, int c, int n, int tid, int nthreads,
                                                        unsigned int n_ctas, bool *last_s) {
    for (int k = tid; k < 2 * c; k += nthreads) atomicAdd(&bn.ws[k], (double)s_sum[k]);
    __threadfence();
    __syncthreads();
    unsigned int *ticket = reinterpret_cast<unsigned int *>(bn.ws + 2 * c);
    if (tid == 0) *last_s = atomicAdd(ticket, 1u) == n_ctas - 1;
    __syncthreads();
    if (!*last_s) return;
    __threadfence();
    for (int k = tid; k < c; k += nthreads) {
        const double sum = __ldcg(&bn.ws[k]), sq = __ldcg(&bn.ws[c + k]);
        const double m = sum / n;
        double var = sq / n - m * m;
        if (var < 0) var = 0;
        bn.mean[k] = (float)m;
        bn.invstd[k] = (float)(1.0 / sqrt(var + (double)bn.eps));
        if (bn.running_mean) bn.running_mean[k] = (float)((1.0 - bn.momentum) * bn.running_mean[k] + bn.momentum * m);
        if (bn.running_var) {
            const double unbiased = n > 1 ? var * n / (n - 1.0) : var;
            bn.running_var[k] = (float)((1.0 - bn.momentum) * bn.running_var[k] + bn.momentum * unbiased);
        }
        bn.ws[k] = 0.0;
        bn.ws[c + k] = 0.0;
    }
    if (tid == 0) {
        *ticket = 0u;
        if (bn.num_batches_tracked) *bn.num_batches_tracked += 1;
    }
}

struct Params {
    const __nv_bfloat16 *x_hi, *x_lo;  // [n_in, cin] planes
    const int32_t *nbr;
    int n_rows, kvol, n_tiles, n_super, T;
    const uint8_t *wpack;
    int cin, cout, nchunks;
    const float *bias;
    const int32_t *out_rows;
    float *y;
    int ldy, accumulate;
    const uint32_t *tile_mask;
    int a_slots, b_slots, b_plane, acc_cols;
    int dbl;                              // 1: two accumulator sets, the epilogue of a tile group overlaps the next group's MMAs
    float *ws;                            // split mode: partial tiles [ksplit][n_rows][cout]
    const int32_t *part;                  // range mode: tiles [part[b], part[b + 1]) belong to CTA b (NULL: equal counts)
    int ksplit, n_units;                  // small maps: the offsets of a tile are dealt to `ksplit` CTAs (T == 1)
    uint32_t part_mask[US3D_MAX_KVOL];    // offsets handled by part q (k % ksplit == q)
    long long *prof;  // optional per-CTA wait-cycle counters of the MMA thread (debug)
    BnFuse bn;        // range mode: statistics of y from the epilogue; split mode: from k_reduce_parts
};

template <int PASSES, int LAG, bool FUSE, bool PROF>
__global__ void __launch_bounds__(THREADS, 1) k_spconv_mt(Params p) {
    static_assert(!FUSE || PASSES == 3, "the fused [W_hi | W_lo] operand exists in three-term mode only");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_full[MAX_A], a_empty[MAX_A], b_full[MAX_B], b_empty[MAX_B], acc_full[2 * MAX_T], acc_empty[2 * MAX_T];  // [set][tile]
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t idx_full[NIDX], idx_empty[NIDX];
    __shared__ int32_t idx_ring[NIDX][MAX_T][M];
    __shared__ float bn_sum[2 * 256];  // column sums / sums of squares of this CTA's output rows (cout <= 256)
    __shared__ bool bn_last;
    const bool bn_on = p.bn.ws != nullptr && p.ksplit == 1;
    constexpr int NPL = PASSES == 3 ? 2 : 1;
    constexpr int A_SLOT = NPL * A_PLANE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_base = smem_u32(smem);
    const uint32_t b_base = a_base + (uint32_t)p.a_slots * A_SLOT;
    const int b_slot_bytes = NPL * p.b_plane;
    const uint32_t all_k = p.kvol >= 32 ? 0xFFFFFFFFu : ((1u << p.kvol) - 1u);
    const int T = p.T;
    uint32_t tmem_cols = 32;  // tcgen05.alloc takes a power of two >= 32 (T = 3 accumulators of 128 columns -> 512)
    while (tmem_cols < (uint32_t)(T * p.acc_cols * (1 + p.dbl))) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < p.a_slots; ++s) {
            mbar_init(smem_u32(&a_full[s]), LAG == 0 ? PROD_WARPS * 32 : PROD_WARPS);
            mbar_init(smem_u32(&a_empty[s]), 1);
        }
        for (int s = 0; s < p.b_slots; ++s) {
            mbar_init(smem_u32(&b_full[s]), 1);
            mbar_init(smem_u32(&b_empty[s]), T);  // every MMA warp releases the slab
        }
        for (int s = 0; s < NIDX; ++s) {
            mbar_init(smem_u32(&idx_full[s]), 32);
            mbar_init(smem_u32(&idx_empty[s]), PROD_WARPS);
        }
        for (int t = 0; t < 2 * MAX_T; ++t) {  // per accumulator: MMA warp t -> the epilogue set of tile t -> MMA warp t
            mbar_init(smem_u32(&acc_full[t]), 1);
            mbar_init(smem_u32(&acc_empty[t]), 4);
        }
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, tmem_cols);
    if (bn_on && threadIdx.x < 2 * 256) bn_sum[threadIdx.x] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait();     // everything above overlaps the previous kernel's tail; nothing below runs before its writes are visible
    pdl_trigger();  // the next kernel of the stream may set itself up while this one works

    auto tile_kmask = [&](int tile, uint32_t pm) -> uint32_t {
        if (tile >= p.n_tiles) return 0u;
        return (p.tile_mask ? (p.tile_mask[tile] & all_k) : all_k) & pm;
    };

    // Work of this CTA.  Range mode (ksplit == 1): a contiguous range of output tiles — equal cost per CTA when the caller
    // supplies a partition (cost of a tile = its active offsets), equal counts otherwise — walked in groups of up to T tiles
    // that share every weight slab.  Split mode (small maps): unit u = (tile u / ksplit, offsets k % ksplit == u % ksplit),
    // dealt round-robin; partial sums meet in y through vector reds.
    int r_begin = 0, r_end = 0;
    if (p.ksplit == 1) {
        if (p.part != nullptr) {
            r_begin = p.part[blockIdx.x];
            r_end = p.part[blockIdx.x + 1];
        } else {
            r_begin = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x);
            r_end = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);
        }
    }
    auto unit = [&](int i, int &tile0, int &nt, uint32_t &pm) -> bool {
        if (p.ksplit == 1) {
            tile0 = r_begin + i * T;
            if (tile0 >= r_end) return false;
            nt = min(T, r_end - tile0);
            pm = 0xFFFFFFFFu;
            return true;
        }
        const int u = blockIdx.x + i * gridDim.x;
        if (u >= p.n_units) return false;
        tile0 = u / p.ksplit;
        nt = 1;
        pm = p.part_mask[u - tile0 * p.ksplit];
        return true;
    };

    if (warp < PROD_WARPS) {
        // ------------------------------------------------------------------ A producers
        // Neighbour indices come from the index ring in shared memory (filled ahead of time by the index warp): no global
        // load sits between two gathers, and no index lives in a register across an item.
        constexpr int RSTEP = PROD_WARPS * 4, RPT = M / RSTEP;  // rows per thread: rbase + RSTEP i
        const int grp = tid & 7;     // 16-byte chunk within the 128-byte row
        const int rbase = tid >> 3;
        int item = 0, signalled = 0;
        // Ring slots are dealt to the tiles of a group: tile t owns slots t, t + T, t + 2T, ... (a_slots is a multiple of T), so a
        // slot is always consumed by the same MMA warp, which therefore observes every phase of its barrier in turn (a parity wait
        // must not lag or lead its barrier by more than one phase).
        const int spt = p.a_slots / T;  // slots per tile
        const uint32_t row_bytes = (uint32_t)p.cin * 2u;
        int pos[MAX_T] = {0, 0, 0, 0};
        uint32_t par[MAX_T] = {0, 0, 0, 0};
        int hist[4] = {0, 0, 0, 0};  // slots of the last items (wait_group variants signal LAG items late)
        int is = 0;          // index-ring stage of the current (super-tile, offset)
        uint32_t ipar = 0;
        int tile0, nt;
        uint32_t pm;
        for (int ui = 0; unit(ui, tile0, nt, pm); ++ui) {
            uint32_t m[MAX_T], U = 0;
#pragma unroll
            for (int t = 0; t < MAX_T; ++t) {
                m[t] = t < nt ? tile_kmask(tile0 + t, pm) : 0u;
                U |= m[t];
            }
            for (int k = 0; k < p.kvol; ++k) {
                if (!((U >> k) & 1u)) continue;
                mbar_wait(smem_u32(&idx_full[is]), ipar, 6);
                for (int c = 0; c < p.nchunks; ++c) {
                    const int c0 = c * KC + grp * 8;
                    const bool col_ok = c0 < p.cin;
                    // byte address of this thread's 16-byte column chunk in row 0 of each plane; a row is one 32 x 32 -> 64 bit
                    // multiply-add away (IMAD.WIDE.U32), absent neighbours go through the copy's ignore-src predicate
                    const uint8_t *col_hi = reinterpret_cast<const uint8_t *>(p.x_hi) + (col_ok ? c0 * 2 : 0);
                    const uint8_t *col_lo = reinterpret_cast<const uint8_t *>(p.x_lo) + (col_ok ? c0 * 2 : 0);
#pragma unroll
                    for (int t = 0; t < MAX_T; ++t) {
                        if (!((m[t] >> k) & 1u)) continue;
                        const int lim = col_ok ? p.n_rows - (tile0 + t) * M : 0;  // rows of the tile inside the map (0: nothing to fetch)
                        int idx[RPT];
#pragma unroll
                        for (int i = 0; i < RPT; ++i) idx[i] = idx_ring[is][t][rbase + RSTEP * i];
                        const int ps = t + T * pos[t];
                        const uint32_t ppar = par[t];
                        if (++pos[t] == spt) {
                            pos[t] = 0;
                            par[t] ^= 1u;
                        }
                        mbar_wait(smem_u32(&a_empty[ps]), ppar ^ 1, 0);
                        const uint32_t slot = a_base + (uint32_t)ps * A_SLOT + (uint32_t)rbase * 128u + (uint32_t)((grp ^ (rbase & 7)) << 4);
#pragma unroll
                        for (int i = 0; i < RPT; ++i) {
                            const bool ign = idx[i] < 0 || rbase + RSTEP * i >= lim;
                            const uint64_t off = (uint64_t)(uint32_t)max(idx[i], 0) * row_bytes;
                            cp_async16_pred(slot + (uint32_t)i * (128u * RSTEP), col_hi + off, ign);
                            if (PASSES == 3) cp_async16_pred(slot + (uint32_t)i * (128u * RSTEP) + A_PLANE, col_lo + off, ign);
                        }
                        if (LAG == 0) {
                            cp_async_arrive_noinc(smem_u32(&a_full[ps]));
                        } else {
                            hist[item & 3] = ps;
                            cp_async_commit();
                            if (item - signalled >= LAG) {
                                cp_async_wait<(LAG > 0 ? LAG : 1)>();
                                fence_proxy_async();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(smem_u32(&a_full[hist[signalled & 3]]));
                                ++signalled;
                            }
                        }
                        ++item;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&idx_empty[is]));
                if (++is == NIDX) {
                    is = 0;
                    ipar ^= 1u;
                }
            }
        }
        cp_async_wait<0>();
        if (LAG != 0) {
            fence_proxy_async();
            __syncwarp();
            for (; signalled < item; ++signalled)
                if (lane == 0) mbar_arrive(smem_u32(&a_full[hist[signalled & 3]]));
        }
    } else if (warp == IDX_WARP) {
        // ------------------------------------------------------------------ neighbour indices, NIDX offsets ahead
        // 4-byte cp.async straight from the kernel map into the index ring; completion by mbarrier (noinc): the warp never
        // waits for a load.  Rows past the map and tiles without the offset are never read by the producers.
        int is = 0;
        uint32_t ipar = 0;
        int tile0, nt;
        uint32_t pm;
        for (int ui = 0; unit(ui, tile0, nt, pm); ++ui) {
            uint32_t m[MAX_T], U = 0;
#pragma unroll
            for (int t = 0; t < MAX_T; ++t) {
                m[t] = t < nt ? tile_kmask(tile0 + t, pm) : 0u;
                U |= m[t];
            }
            for (int k = 0; k < p.kvol; ++k) {
                if (!((U >> k) & 1u)) continue;
                mbar_wait(smem_u32(&idx_empty[is]), ipar ^ 1, 7);
                const int32_t *src_k = p.nbr + (size_t)k * p.n_rows;
#pragma unroll
                for (int t = 0; t < MAX_T; ++t) {
                    if (!((m[t] >> k) & 1u)) continue;
                    const int row0 = (tile0 + t) * M;
#pragma unroll
                    for (int q = 0; q < M / 32; ++q) {
                        const int r = lane + 32 * q;
                        const bool ok = row0 + r < p.n_rows;
                        cp_async4_pred(smem_u32(&idx_ring[is][t][r]), src_k + (ok ? row0 + r : 0), !ok);
                    }
                }
                cp_async_arrive_noinc(smem_u32(&idx_full[is]));
                if (++is == NIDX) {
                    is = 0;
                    ipar ^= 1u;
                }
            }
        }
        cp_async_wait_all();
    } else if (warp == B_WARP) {
        // ------------------------------------------------------------------ weight slabs
        if (lane == 0) {
            int bitem = 0;
            int tile0, nt;
            uint32_t pm;
            for (int ui = 0; unit(ui, tile0, nt, pm); ++ui) {
                uint32_t U = 0;
                for (int t = 0; t < nt; ++t) U |= tile_kmask(tile0 + t, pm);
                for (int k = 0; k < p.kvol; ++k) {
                    if (!((U >> k) & 1u)) continue;
                    for (int c = 0; c < p.nchunks; ++c, ++bitem) {
                        const int s = bitem % p.b_slots;
                        const uint32_t par = (bitem / p.b_slots) & 1;
                        mbar_wait(smem_u32(&b_empty[s]), par ^ 1, 1);
                        const uint32_t bar = smem_u32(&b_full[s]);
                        const uint32_t dst = b_base + (uint32_t)s * b_slot_bytes;
                        const uint8_t *src = p.wpack + ((size_t)k * p.nchunks + c) * (size_t)b_slot_bytes;
                        mbar_arrive_expect_tx(bar, (uint32_t)b_slot_bytes);
                        bulk_g2s(dst, src, (uint32_t)p.b_plane, bar);
                        if (PASSES == 3) bulk_g2s(dst + p.b_plane, src + p.b_plane, (uint32_t)p.b_plane, bar);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP || (warp >= MMA_WARP_EXTRA0 && warp - MMA_WARP_EXTRA0 + 1 < T)) {
        // ------------------------------------------------------------------ MMA issuers, one warp per tile of the group
        // The per-item work of an issuing thread (barrier test, proxy fence, descriptors, 8-12 MMAs, commit, bookkeeping) is a
        // serial instruction stream of ~1000 cycles against ~450 tensor-pipe cycles of MMAs (profiles/r2_conv_pipeline.md): with
        // one issuer per CTA the tensor pipe idles 60 % of the time.  Tile t of a group therefore has its own issuing warp mw = t:
        // every warp walks the same item sequence (ring positions are pure counters), waits / issues / frees only the slots of
        // its own tile, and all of them release the weight slab and the accumulators (barrier counts = T).
        const int mw = warp == MMA_WARP ? 0 : warp - MMA_WARP_EXTRA0 + 1;
        // The whole warp runs the (warp-uniform) control flow so that descriptors and ring positions live in
        // uniform registers; one elected lane issues tcgen05.mma / tcgen05.commit.  (A single-lane branch makes
        // ptxas wrap every UTCHMMA in an elect + R2UR.BROADCAST loop: ~80 issue cycles per MMA, measured.)
        const uint32_t idesc = idesc_bf16(p.cout);
        const uint32_t idesc2 = idesc_bf16(2 * p.cout);  // FUSE: B = [W_hi rows | W_lo rows]
        const uint64_t a_desc0 = desc_k_sw128(a_base), b_desc0 = desc_k_sw128(b_base);
        const int spt = p.a_slots / T;  // ring slots of this warp's tile: mw, mw + T, ... (see the producers)
        int apos = 0, bs = 0, siter = 0, item = 0, bitem = 0;
        uint32_t apar = 0, bpar = 0;
        long long w_acc = 0, w_b = 0, w_a = 0, t_first = 0;
        const long long t_begin = PROF ? clock64() : 0;  // the cycle counters exist only in the profiling instantiation
        int tile0, nt;
        uint32_t pm;
        for (int ui = 0; unit(ui, tile0, nt, pm); ++ui, ++siter) {
            uint32_t m[MAX_T], U = 0;
#pragma unroll
            for (int t = 0; t < MAX_T; ++t) {
                m[t] = t < nt ? tile_kmask(tile0 + t, pm) : 0u;
                U |= m[t];
            }
            long long tw0 = PROF ? clock64() : 0;
            // accumulator set of this group and how often it has been used before (the barrier's phase)
            const int aset = p.dbl ? (siter & 1) : 0, ause = p.dbl ? (siter >> 1) : siter;
            mbar_wait(smem_u32(&acc_empty[aset * MAX_T + mw]), (ause & 1) ^ 1, 2);
            if (PROF) w_acc += clock64() - tw0;
            tc_fence_after();
            uint32_t started = 0;
            for (int k = 0; k < p.kvol; ++k) {
                if (!((U >> k) & 1u)) continue;
                for (int c = 0; c < p.nchunks; ++c) {
                    long long tw1 = PROF ? clock64() : 0;
                    mbar_wait(smem_u32(&b_full[bs]), bpar, 3);
                    if (PROF) w_b += clock64() - tw1;
                    const int ksteps = min(KC, p.cin - c * KC) / 16;
                    const uint64_t db_hi = b_desc0 + (uint64_t)((uint32_t)(bs * b_slot_bytes) >> 4);
                    const uint64_t db_lo = db_hi + (uint64_t)((uint32_t)p.b_plane >> 4);
#pragma unroll
                    for (int t = 0; t < MAX_T; ++t) {
                        if (!((m[t] >> k) & 1u)) continue;
                        if (t != mw) continue;  // another warp's item
                        const int as = mw + T * apos;
                        long long tw2 = PROF ? clock64() : 0;
                        mbar_wait(smem_u32(&a_full[as]), apar, 4);
                        if (PROF) {
                            w_a += clock64() - tw2;
                            if (item == 0) t_first = clock64() - t_begin;
                        }
                        if (LAG == 0) fence_proxy_async();  // cp.async writes (generic proxy) -> tcgen05 reads (async proxy)
                        tc_fence_after();
                        const uint64_t da_hi = a_desc0 + (uint64_t)((uint32_t)(as * A_SLOT) >> 4);
                        const uint64_t da_lo = da_hi + (uint64_t)(A_PLANE >> 4);
                        const uint32_t acc = tmem_base + (uint32_t)((aset * T + t) * p.acc_cols);
                        const uint32_t first = (started >> t) & 1u;
                        if (elect_one()) {
                            for (int kk = 0; kk < ksteps; ++kk) {
                                const uint64_t adv = (uint64_t)(kk * 2);
                                if (FUSE) {
                                    umma(acc, da_hi + adv, db_hi + adv, idesc2, first | (kk != 0));
                                    umma(acc, da_lo + adv, db_hi + adv, idesc, 1);
                                } else {
                                    umma(acc, da_hi + adv, db_hi + adv, idesc, first | (kk != 0));
                                    if (PASSES == 3) {
                                        umma(acc, da_lo + adv, db_hi + adv, idesc, 1);
                                        umma(acc, da_hi + adv, db_lo + adv, idesc, 1);
                                    }
                                }
                            }
                            umma_commit(smem_u32(&a_empty[as]));
                        }
                        __syncwarp();
                        started |= 1u << t;
                        ++item;
                        if (++apos == spt) {
                            apos = 0;
                            apar ^= 1u;
                        }
                    }
                    if (elect_one()) umma_commit(smem_u32(&b_empty[bs]));
                    __syncwarp();
                    ++bitem;
                    if (++bs == p.b_slots) {
                        bs = 0;
                        bpar ^= 1u;
                    }
                }
            }
            if (elect_one()) {
                if (started)
                    umma_commit(smem_u32(&acc_full[aset * MAX_T + mw]));
                else
                    mbar_arrive(smem_u32(&acc_full[aset * MAX_T + mw]));
            }
            __syncwarp();
        }
        if (PROF && p.prof != nullptr && lane == 0 && mw == 0) {
            long long *o = p.prof + (size_t)blockIdx.x * 8;
            o[0] = clock64() - t_begin;
            o[1] = w_acc;
            o[2] = w_b;
            o[3] = w_a;
            o[4] = item;
            o[5] = bitem;
            o[6] = t_first;  // cycles until the first gathered tile has landed (prologue + one index + one gather round trip)
        }
    } else if ((warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) || (warp >= EPI2_WARP0 && warp < EPI2_WARP0 + 4)) {
        // ------------------------------------------------------------------ epilogue
        // Two sets of four warps; set e reads out tiles e, e + 2 of every group.  Each tile has its own accumulator barriers, so the
        // MMA warp of a tile starts the next group as soon as ITS accumulator has been read, while the other tiles are still being
        // written; everything that does not depend on the accumulator (tile mask, output row) is fetched before the wait.
        const int eset = warp >= EPI2_WARP0 ? 1 : 0;
        const int quarter = warp & 3;
        const bool vec = (p.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        int siter = 0;
        const bool split = p.ksplit > 1;  // the offset parts write partial tiles to the workspace; k_reduce_parts sums them
        int tile0, nt;
        uint32_t pm;
        for (int ui = 0; unit(ui, tile0, nt, pm); ++ui, ++siter) {
            const int part = p.ksplit == 1 ? 0 : (blockIdx.x + ui * gridDim.x) % p.ksplit;
            const int aset = p.dbl ? (siter & 1) : 0, ause = p.dbl ? (siter >> 1) : siter;
            for (int t = eset; t < T; t += EPI_SETS) {
                const int tile = tile0 + t;
                const bool live = t < nt && tile < p.n_tiles;
                const bool has_acc = live && tile_kmask(tile, pm) != 0;
                const int j = tile * M + quarter * 32 + lane;
                const bool row_ok = live && j < p.n_rows;
                float *yrow = nullptr;
                if (row_ok)
                    yrow = split ? p.ws + ((size_t)part * p.n_rows + j) * p.cout
                                 : p.y + (size_t)(p.out_rows ? p.out_rows[j] : j) * p.ldy;
                mbar_wait(smem_u32(&acc_full[aset * MAX_T + t]), ause & 1, 5);
                tc_fence_after();
                const uint32_t acc_addr = tmem_base + (uint32_t)((aset * T + t) * p.acc_cols) + ((uint32_t)(quarter * 32) << 16);
                for (int col = 0; live && col < p.cout; col += 16) {
                    float acc[16];
                    if (has_acc) {
                        if (FUSE) {  // both column groups requested before the one wait
                            float acc2[16];
                            tmem_ld16_nowait(acc_addr + (uint32_t)col, acc);
                            tmem_ld16_nowait(acc_addr + (uint32_t)(p.cout + col), acc2);
                            tmem_ld_wait(acc, acc2);
#pragma unroll
                            for (int e = 0; e < 16; ++e) acc[e] += acc2[e];
                        } else {
                            tmem_ld16(acc_addr + (uint32_t)col, acc);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) acc[e] = 0.f;
                    }
                    if (split) {  // plain stores of the partial tile (cout % 16 == 0, workspace 16-byte aligned)
                        if (!row_ok) continue;
#pragma unroll
                        for (int e = 0; e < 16; e += 4)
                            *reinterpret_cast<float4 *>(yrow + col + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
                        continue;
                    }
                    if (p.bias)
#pragma unroll
                        for (int e = 0; e < 16; ++e) acc[e] += __ldg(p.bias + col + e);
                    if (row_ok) {
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
                    if (bn_on) {  // the whole warp: column sums of the 32 rows, one shared-memory add per column
                        float sq[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            acc[e] = row_ok ? acc[e] : 0.f;
                            sq[e] = acc[e] * acc[e];
                        }
                        warp_colsum16(acc, lane);
                        warp_colsum16(sq, lane);
                        if (!(lane & 1)) {
                            const int cc = col + warp_colsum16_col(lane);
                            atomicAdd(&bn_sum[cc], acc[0]);
                            atomicAdd(&bn_sum[p.cout + cc], sq[0]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&acc_empty[aset * MAX_T + t]));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, tmem_cols);
    if (bn_on) bn_publish_and_finalize(p.bn, bn_sum, p.cout, p.n_rows, tid, THREADS, gridDim.x, &bn_last);
}

// Cost-weighted partition of the output tiles over G CTAs: tile i costs (active offsets of its mask) + 2 (index / epilogue
// overhead); part[b] = first tile of CTA b, chosen so that every CTA's range carries ~1/G of the total.  One block.
__global__ void __launch_bounds__(1024) k_partition(const uint32_t *__restrict__ tile_mask, int n_tiles, int kvol, int G,
                                                    int32_t *__restrict__ part) {
    __shared__ long long wsum[32];
    __shared__ long long carry_s, total_s;
    const uint32_t all_k = kvol >= 32 ? 0xFFFFFFFFu : ((1u << kvol) - 1u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto cost = [&](int i) -> long long { return i < n_tiles ? (long long)__popc(tile_mask[i] & all_k) + 2 : 0; };
    // pass 1: total
    long long acc = 0;
    for (int i = tid; i < n_tiles; i += blockDim.x) acc += cost(i);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) wsum[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int w = 0; w < 32; ++w) t += wsum[w];
        total_s = t;
        carry_s = 0;
    }
    __syncthreads();
    const long long total = total_s;
    if (tid <= G) {
        if (tid == 0) part[0] = 0;
        if (tid == G) part[G] = n_tiles;
    }
    // pass 2: running prefix, chunk by chunk; tile i opens CTA b's range when the prefix BEFORE it crosses b * total / G
    for (int base = 0; base < n_tiles; base += blockDim.x) {
        const int i = base + tid;
        long long c = cost(i), incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        long long woff = 0;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        const long long before = carry_s + woff + incl - c;  // cost of tiles [0, i)
        if (i < n_tiles && total > 0) {
            // boundaries b with  before < b * total / G <= before + c   start AFTER tile i (tile i closes CTA b - 1's range)
            long long b_lo = (before * G) / total + 1, b_hi = ((before + c) * G) / total;
            for (long long b = b_lo; b <= b_hi && b < G; ++b)
                if (b >= 1) part[b] = i + 1;
        }
        __syncthreads();
        if (tid == blockDim.x - 1) carry_s = carry_s + woff + incl;
        __syncthreads();
    }
}

// Split mode, second half: y[row] (=|+=) bias + sum_q ws[q][row], parts added in index order (bit-reproducible; the vector
// reds this replaces cost ~25 us per 128-row tile on the B200: ~76 G fp32 atomics/s chip-wide).
__global__ void __launch_bounds__(256) k_reduce_parts(const float *__restrict__ ws, int ksplit, int n_rows, int cout,
                                                      const float *__restrict__ bias, float *__restrict__ y, int ldy, int accumulate,
                                                      BnFuse bn) {
    __shared__ float bn_sum[2 * 256];
    __shared__ bool bn_last;
    const bool bn_on = bn.ws != nullptr;
    if (bn_on) {
        for (int k = threadIdx.x; k < 2 * cout; k += blockDim.x) bn_sum[k] = 0.f;
        __syncthreads();
    }
    pdl_wait();
    pdl_trigger();
    const int g = cout / 4;
    const long long total = (long long)n_rows * g;
    const size_t plane = (size_t)n_rows * cout;
    // the launcher makes the grid stride a multiple of g, so a thread stays on ONE column group: its sums live in registers
    float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
    int my_c = -1;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / g), c = (int)(e - (long long)r * g) * 4;
        float4 acc = *reinterpret_cast<const float4 *>(ws + (size_t)r * cout + c);
        for (int q = 1; q < ksplit; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(ws + q * plane + (size_t)r * cout + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (bias != nullptr) {
            acc.x += bias[c]; acc.y += bias[c + 1]; acc.z += bias[c + 2]; acc.w += bias[c + 3];
        }
        float *dst = y + (size_t)r * ldy + c;
        if (accumulate) {
            acc.x += dst[0]; acc.y += dst[1]; acc.z += dst[2]; acc.w += dst[3];
        }
        dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w;
        if (bn_on) {
            my_c = c;
            s4[0] += acc.x; s4[1] += acc.y; s4[2] += acc.z; s4[3] += acc.w;
            q4[0] += acc.x * acc.x; q4[1] += acc.y * acc.y; q4[2] += acc.z * acc.z; q4[3] += acc.w * acc.w;
        }
    }
    if (bn_on && my_c >= 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            atomicAdd(&bn_sum[my_c + i], s4[i]);
            atomicAdd(&bn_sum[cout + my_c + i], q4[i]);
        }
    }
    if (bn_on) {
        __syncthreads();
        bn_publish_and_finalize(bn, bn_sum, cout, n_rows, threadIdx.x, blockDim.x, gridDim.x, &bn_last);
    }
}

template <int PASSES, bool FUSE>
static cudaError_t launch(int lag, int grid, size_t smem, cudaStream_t st, const Params &p) {
    // every instantiation needs the opt-in for > 48 KB of dynamic shared memory once per device
    static bool attr_done[8][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    auto go = [&](auto kernel, int slot) -> cudaError_t {
        if (dev >= 0 && dev < 64 && !attr_done[slot][dev]) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 214 * 1024);  // + 11 KB static (index ring, barriers, BatchNorm sums) <= 227 KB
            if (e != cudaSuccess) return e;
            attr_done[slot][dev] = true;
        }
        return launch_pdl(kernel, dim3(grid), dim3(THREADS), smem, st, p);
    };
    if (p.prof != nullptr) {  // the instantiation with the MMA thread's cycle counters (us3d_debug_set_prof)
        if (lag <= 0) return go(k_spconv_mt<PASSES, 0, FUSE, true>, 4);
        if (lag == 1) return go(k_spconv_mt<PASSES, 1, FUSE, true>, 5);
        if (lag == 2) return go(k_spconv_mt<PASSES, 2, FUSE, true>, 6);
        return go(k_spconv_mt<PASSES, 3, FUSE, true>, 7);
    }
    if (lag <= 0) return go(k_spconv_mt<PASSES, 0, FUSE, false>, 0);
    if (lag == 1) return go(k_spconv_mt<PASSES, 1, FUSE, false>, 1);
    if (lag == 2) return go(k_spconv_mt<PASSES, 2, FUSE, false>, 2);
    return go(k_spconv_mt<PASSES, 3, FUSE, false>, 3);
}

}  // namespace mt
}  // namespace us3d

using namespace us3d;

static long long *g_prof = nullptr;
// launcher defaults (set from the measurements in profiles/r2_conv_tuning.md)
static bool g_default_fuse = true;
static int g_default_dbl = 0;
static int g_default_lag = 0;  // 0: completion by cp.async.mbarrier.arrive.noinc (measured best on every level, profiles/r2_conv_tuning.md); -1: wait_group look-ahead 1 (three-term) / 2 (single pass)

static int g_tune_a_slots = 0, g_tune_lag = 0, g_tune_T = 0, g_tune_fuse = 0, g_tune_dbl = 0;

extern "C" {

/* debug / tuning hooks (not part of the drop-in surface): per-CTA wait counters of the MMA thread, ring overrides */
void us3d_debug_set_prof(void *buf) { g_prof = (long long *)buf; }
/* a_slots / T: 0 = launcher's choice.  lag: 0 = launcher's choice, 1..3 = cp.async.wait_group look-ahead, 9 = completion by
   cp.async.mbarrier.arrive.noinc (producers never wait).  fuse: 0 = launcher's choice, 1 = off, 2 = on (where eligible). */
void us3d_debug_set_tuning4(int a_slots, int lag, int T, int fuse) {
    g_tune_a_slots = a_slots;
    g_tune_lag = lag;
    g_tune_T = T;
    g_tune_fuse = fuse;
}
void us3d_debug_set_tuning(int a_slots, int lag, int T) { us3d_debug_set_tuning4(a_slots, lag, T, 0); }
/* accumulator sets: 0 = launcher's choice, 1 = one, 2 = two (where TMEM has room) */
void us3d_debug_set_tuning_acc(int sets) { g_tune_dbl = sets; }

int us3d_spconv_partition(const uint32_t *tile_mask, int n_tiles, int kvol, int32_t *partition, void *stream_) {
    US3D_CHECK_ARG(tile_mask != nullptr && partition != nullptr && n_tiles >= 0 && kvol >= 1 && kvol <= US3D_MAX_KVOL, "spconv_partition: bad arguments");
    mt::k_partition<<<1, 1024, 0, (cudaStream_t)stream_>>>(tile_mask, n_tiles, kvol, num_sms(), partition);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_spconv_partition_size(void) { return num_sms() + 1; }

int us3d_spconv_gather_mt(const void *x_hi, const void *x_lo, int n_in, const int32_t *nbr, int n_rows, int kvol,
                          const void *wpack, int cin, int cout, int passes, const float *bias, const int32_t *out_rows,
                          float *y, int ldy, int accumulate, const uint32_t *tile_mask, const int32_t *partition, void *workspace,
                          long long workspace_bytes, void *stream_) {
    return us3d_spconv_gather_mt_bn(x_hi, x_lo, n_in, nbr, n_rows, kvol, wpack, cin, cout, passes, bias, out_rows, y, ldy, accumulate,
                                    tile_mask, partition, workspace, workspace_bytes, nullptr, stream_);
}

int us3d_spconv_gather_mt_bn(const void *x_hi, const void *x_lo, int n_in, const int32_t *nbr, int n_rows, int kvol,
                             const void *wpack, int cin, int cout, int passes, const float *bias, const int32_t *out_rows,
                             float *y, int ldy, int accumulate, const uint32_t *tile_mask, const int32_t *partition, void *workspace,
                             long long workspace_bytes, const us3d_bn_fuse_t *bn, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "spconv_gather_mt: kvol %d out of range", kvol);
    US3D_CHECK_ARG(passes == 1 || passes == 3, "spconv_gather_mt: passes must be 1 or 3");
    US3D_CHECK_ARG(us3d_spconv_tc_supported(cin, cout), "spconv_gather_mt: unsupported channel counts %d -> %d", cin, cout);
    US3D_CHECK_ARG(x_hi != nullptr && (passes == 1 || x_lo != nullptr), "spconv_gather_mt: missing activation plane");
    US3D_CHECK_ARG((reinterpret_cast<uintptr_t>(x_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_lo) & 15) == 0,
                   "spconv_gather_mt: planes must be 16-byte aligned");
    US3D_CHECK_ARG(ldy >= cout && n_in > 0, "spconv_gather_mt: bad sizes");
    US3D_CHECK_ARG(bn == nullptr || (bn->ws != nullptr && bn->mean != nullptr && bn->invstd != nullptr && !accumulate && n_rows > 0),
                   "spconv_gather_mt_bn: statistics need ws / mean / invstd, a non-empty map and accumulate == 0");
    if (n_rows == 0) return 0;
    mt::Params p;
    p.bn = mt::BnFuse{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f};
    if (bn != nullptr)
        p.bn = mt::BnFuse{bn->ws, bn->mean, bn->invstd, bn->running_mean, bn->running_var, bn->num_batches_tracked, bn->eps, bn->momentum};
    p.x_hi = (const __nv_bfloat16 *)x_hi; p.x_lo = (const __nv_bfloat16 *)x_lo;
    p.nbr = nbr; p.n_rows = n_rows; p.kvol = kvol; p.n_tiles = ceil_div(n_rows, mt::M);
    p.wpack = (const uint8_t *)wpack; p.cin = cin; p.cout = cout; p.nchunks = ceil_div(cin, mt::KC);
    p.bias = bias; p.out_rows = out_rows; p.y = y; p.ldy = ldy; p.accumulate = accumulate; p.tile_mask = tile_mask;
    const int npl = passes == 3 ? 2 : 1;
    const int sms = num_sms();
    p.prof = g_prof;
    // Work decomposition.  The kernel's cost is ~ the (tile, offset) products on the busiest CTA plus ~2 of them per tile for the
    // epilogue.  With the offsets of a tile dealt to `ks` CTAs (each writes a partial tile to the workspace, k_reduce_parts sums
    // them in a second launch) the busiest CTA runs ceil(tiles * ks / SMs) units of ceil(kvol / ks) offsets.  ks = 1 needs no
    // second pass and shares weight slabs between the tiles of a CTA, so a split has to win by 25 %.  Without a workspace the
    // map is never split.  (Round 1 split with vector reds into y: measured ~25 us per 128-row tile on the B200.)
    int ksplit = 1;
    if (out_rows == nullptr && kvol > 1 && workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0) {
        long long best = (long long)ceil_div(p.n_tiles, sms) * (kvol + 2) * 100;
        for (int ks = 2; ks <= kvol; ++ks) {
            if ((long long)ks * n_rows * cout * (long long)sizeof(float) > workspace_bytes) break;
            const long long c = (long long)ceil_div((long long)p.n_tiles * ks, sms) * (ceil_div(kvol, ks) + 2) * 125;
            if (c < best) {
                best = c;
                ksplit = ks;
            }
        }
    }
    // fused [W_hi | W_lo] operand: one N = 2 Cout MMA for the two products that share X_hi (three-term mode, N <= 256).
    // Two accumulator sets (the epilogue of a tile group overlaps the MMAs of the next) where the epilogue also carries the
    // BatchNorm statistics of a large map: measured on B200 (scripts/tune_acc.py, 200k voxels 96 -> 96 with statistics) fused /
    // one set 291 us, unfused / two sets 270 us, fused / two sets with single-tile groups 365 us; without statistics 262 / 262 us.
    bool fuse = passes == 3 && 2 * cout <= 256 && g_default_fuse;
    int dbl = (bn != nullptr && ksplit == 1) ? 1 : g_default_dbl;
    if (g_tune_dbl == 1) dbl = 0;
    if (g_tune_dbl == 2) dbl = 1;
    if (dbl && 2 * cout > 128 && g_tune_fuse != 2) fuse = false;  // 2 x 256 columns would leave one tile per group
    if (g_tune_fuse == 1) fuse = false;
    if (g_tune_fuse == 2) fuse = passes == 3 && 2 * cout <= 256;
    int cols = 32;
    while (cols < (fuse ? 2 * cout : cout)) cols <<= 1;
    p.acc_cols = cols;
    if (2 * cols > 512) dbl = 0;
    p.dbl = dbl;
    int T = 512 / (cols * (1 + dbl));
    if (T > mt::MAX_T) T = mt::MAX_T;
    p.ws = (float *)workspace;
    if (ksplit > 1) T = 1;
    // range mode: every CTA owns a contiguous range of tiles; no more accumulators than its share of the tiles
    if (ksplit == 1) {
        const int per_cta = ceil_div(p.n_tiles, sms);
        if (T > per_cta) T = per_cta;
    }
    if (g_tune_T > 0 && g_tune_T <= T) T = g_tune_T;
    p.T = T;
    p.n_super = ceil_div(p.n_tiles, T);
    p.ksplit = ksplit;
    p.n_units = ksplit == 1 ? p.n_tiles : p.n_tiles * ksplit;
    // the cost-weighted partition (us3d_spconv_partition) is laid out for one CTA per SM
    p.part = (ksplit == 1 && partition != nullptr && p.n_tiles >= 2 * sms) ? partition : nullptr;
    for (int q = 0; q < US3D_MAX_KVOL; ++q) {
        uint32_t pmask = 0;
        for (int k = q; k < kvol && q < ksplit; k += ksplit) pmask |= 1u << k;
        p.part_mask[q] = ksplit == 1 ? 0xFFFFFFFFu : pmask;
    }
    p.b_plane = cout * 128;
    const int a_slot = npl * mt::A_PLANE, b_slot = npl * p.b_plane;
    const int budget = 208 * 1024;
    p.b_slots = 2;
    p.a_slots = (budget - p.b_slots * b_slot) / a_slot;
    if (p.a_slots > mt::MAX_A) p.a_slots = mt::MAX_A;
    US3D_CHECK_ARG(p.a_slots >= 2, "spconv_gather_mt: operand slots do not fit in shared memory (cout %d)", cout);
    if (p.a_slots == mt::MAX_A && budget - p.a_slots * a_slot - 3 * b_slot >= 0) p.b_slots = 3;
    if (g_tune_a_slots >= 2 && g_tune_a_slots <= p.a_slots) p.a_slots = g_tune_a_slots;
    // the ring slots are dealt to the tiles of a group in equal shares
    if (p.a_slots < T) {
        T = p.a_slots;
        p.T = T;
        p.n_super = ceil_div(p.n_tiles, T);
    }
    p.a_slots = (p.a_slots / T) * T;
    const size_t smem = (size_t)p.a_slots * a_slot + (size_t)p.b_slots * b_slot + 1024;
    const int grid = p.n_units < sms ? p.n_units : sms;
    // Groups a producer warp keeps in flight before it waits for the oldest.  Measured on B200 (200k voxels, 128 -> 96):
    // three-term mode 0.521 ms at lag 3, 0.472 ms at lag 1 — with 32 KB slots a deep lag leaves the MMA warp no landed
    // slot to run ahead on; single-pass mode is best at lag 2.
    int lag = g_default_lag >= 0 ? g_default_lag : (passes == 3 ? 1 : 2);
    if (lag > p.a_slots - 1) lag = p.a_slots - 1;
    if (g_tune_lag >= 1 && g_tune_lag <= p.a_slots - 1) lag = g_tune_lag;
    if (g_tune_lag == 9) lag = 0;
    {
        ProfScope prof(st, 0, n_in, n_rows, kvol, cin, cout);
        cudaError_t e;
        if (passes == 3)
            e = fuse ? mt::launch<3, true>(lag, grid, smem, st, p) : mt::launch<3, false>(lag, grid, smem, st, p);
        else
            e = mt::launch<1, false>(lag, grid, smem, st, p);
        US3D_CUDA(e);
        US3D_LAUNCH_CHECK();
        if (ksplit > 1) {
            long long blocks = ((long long)n_rows * (cout / 4) + 255) / 256;
            const long long cap = p.bn.ws != nullptr ? (long long)sms : (long long)sms * 8;  // with statistics: every block ends in 2 cout fp64 atomics + a ticket
            if (blocks > cap) blocks = cap;
            // grid stride (256 blocks) a multiple of cout / 4 = 4 (cout / 16): blocks a multiple of the odd part of cout / 16
            int odd = cout / 16;
            while (odd % 2 == 0) odd /= 2;
            blocks = (blocks + odd - 1) / odd * odd;
            US3D_CUDA(launch_pdl(mt::k_reduce_parts, dim3((int)blocks), dim3(256), 0, st, (const float *)p.ws, ksplit, n_rows, cout, bias, y, ldy, accumulate, p.bn));
            US3D_LAUNCH_CHECK();
        }
    }
    return 0;
}

/* bytes of workspace that let us3d_spconv_gather_mt split a small map's offsets over all SMs (0: the map is never split) */
long long us3d_spconv_gather_mt_workspace_bytes(int n_rows, int kvol, int cout) {
    const int n_tiles = ceil_div(n_rows, mt::M);
    if (kvol <= 1 || n_tiles >= 2 * num_sms()) return 0;
    return (long long)kvol * n_rows * cout * (long long)sizeof(float);
}

}  // extern "C"
