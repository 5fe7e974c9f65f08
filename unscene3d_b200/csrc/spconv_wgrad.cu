// Sparse-convolution weight gradient on tcgen05 — production kernel.
//
//     dW[k][ci][co] += sum_j X[nbr[k, j]][ci] * dY[j][co]
//
// D[M = ci (128 TMEM lanes)][N = co] accumulates over K = rows j.  Operands are read from the bf16 planes
// (hi, + lo for the three-term split) exactly as they lie in HBM — rows j, channels contiguous — which makes
// them MN-major UMMA operands: every 64-channel block of a 64-row stage is a [64 rows][128 B] slab in the
// SWIZZLE_128B MN-major canonical layout (8-row groups 1024 B apart = SBO, blocks one slab apart = LBO); one
// K = 16 step is two 8-row groups (2048 B).
//
// One CTA = (group of KG = 512 / npad (<= 4) kernel offsets, 128-input-channel block, row split).  For every
// 64-row block of its split the dY tile is fetched ONCE (B ring) and multiplied against the gathered X tile of
// each offset of the group (A ring), each offset accumulating into its own TMEM accumulator — the dY stream,
// which a one-offset-per-CTA layout re-reads 27 times, is read ceil(27 / KG) times.
//
//   warps 0-7  producers: 16-byte cp.async (ignore-src predicate) into the swizzled slabs (zero-fill for absent neighbours / rows
//              past the split); one A ring slot = one plane of one (row block, offset) = 16 KB; neighbour indices are fetched one
//              row block ahead.  Completion is signalled by cp.async.mbarrier.arrive.noinc (the producers never wait for their
//              own copies); the MMA warps cross the generic -> async proxy with fence.proxy.async after their wait
//   warps 8-9  MMA issuers (warp-uniform control flow, one elected lane issues): offset kq of the group belongs to warp kq % 2,
//              with its own TMEM accumulators and its own A ring — tcgen05.mma issue blocks for about the instruction's execution
//              time, so one issuer alone leaves the tensor pipe idle while it waits and books (profiles/r2_wgrad_roles.md)
//   warps 0-7  epilogue at the end: tcgen05.ld -> red.global.add into dW
#include "common.cuh"
#include "tc_common.cuh"

namespace us3d {
namespace wg {

using namespace tcx;

constexpr int R = 64;             // rows (GEMM-K) per stage
constexpr int SLAB = R * 128;     // one 64-channel block of one plane
constexpr int PROD_WARPS = 8;  // a warp sustains one scattered LDGSTS.128 per ~50 cycles: the gather rate scales with the warps
constexpr int MMA_WARP = PROD_WARPS;
constexpr int NMW = 2;  // MMA-issuing warps: offset kq of the group belongs to warp kq % NMW (its own accumulators, its own A ring)
constexpr int THREADS = (PROD_WARPS + NMW) * 32;
constexpr int MAX_A = 6, MAX_B = 3, MAX_KG = 4;  // MAX_A: slots per A ring

struct Params {
    const __nv_bfloat16 *x_hi, *x_lo;    // [n_in, cin]
    const __nv_bfloat16 *dy_hi, *dy_lo;  // [n_rows, cout]
    const int32_t *nbr;
    int n_rows, kvol;
    float *dw;
    int cin, cout, npad, nmma, mblks, splits, rows_per_split, kg, ngroups;
    const uint32_t *tile_mask;
    const int32_t *dy_rows;  // optional: table column j pairs with dY row dy_rows[j] (pattern-ordered tables)
    int a_slots, b_slots, acc_cols;
    long long *prof;  // optional per-CTA cycle counters (profiling instantiation only)
};

template <int PASSES, bool PROF>
__global__ void __launch_bounds__(THREADS, 1) k_wgrad(Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_full[NMW][MAX_A], a_empty[NMW][MAX_A], b_full[MAX_B], b_empty[MAX_B], acc_full;
    __shared__ uint32_t tmem_base_s;
    constexpr int NPL = PASSES == 3 ? 2 : 1;
    constexpr int A_PLANE = 2 * SLAB;  // 128 input channels
    constexpr int A_SLOT = A_PLANE;  // one plane per ring slot
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_plane = (p.npad / 64) * SLAB;
    const int b_slot_bytes = NPL * b_plane;
    const uint32_t a_base = smem_u32(smem);
    const uint32_t b_base = a_base + (uint32_t)(NMW * p.a_slots) * A_SLOT;  // ring m = slots [m a_slots, (m + 1) a_slots)

    int b = blockIdx.x;
    const int split = b % p.splits;
    b /= p.splits;
    const int mblk = b % p.mblks;
    const int grp = b / p.mblks;
    const int k0 = grp * p.kg;
    const int nk = min(p.kg, p.kvol - k0);
    const int r_begin = split * p.rows_per_split;
    const int r_end = min(p.n_rows, r_begin + p.rows_per_split);
    const int ci0 = mblk * 128;
    const uint32_t gmask = ((nk >= 32 ? 0xFFFFFFFFu : ((1u << nk) - 1u)) << k0);

    // offsets of this group with a neighbour somewhere in the 128-row tile holding row r0
    auto active = [&](int r0) -> uint32_t {
        return (p.tile_mask ? p.tile_mask[r0 >> 7] : 0xFFFFFFFFu) & gmask;
    };

    if (tid == 0) {
        for (int m = 0; m < NMW; ++m)
            for (int s = 0; s < p.a_slots; ++s) {
                mbar_init(smem_u32(&a_full[m][s]), PROD_WARPS * 32);
                mbar_init(smem_u32(&a_empty[m][s]), 1);
            }
        for (int s = 0; s < p.b_slots; ++s) {
            mbar_init(smem_u32(&b_full[s]), PROD_WARPS * 32);
            mbar_init(smem_u32(&b_empty[s]), NMW);  // every MMA warp releases the dY tile
        }
        mbar_init(smem_u32(&acc_full), NMW);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, (uint32_t)p.acc_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait();     // (common.cuh) the prologue above overlaps the previous kernel's tail
    pdl_trigger();
    uint32_t touched = 0;  // offsets (relative to k0) that received at least one MMA — same in every role

    if (warp < PROD_WARPS) {
        // ------------------------------------------------------------------ producers
        const int GB = p.cout / 8;  // 16-byte chunks per dY row
        int as[NMW] = {0, 0}, bs = 0;
        uint32_t apar[NMW] = {0, 0}, bpar = 0;
        // this thread's APT (row, 16-byte chunk) cells of an A slot: row = RSTEP i + tid / 16, chunk g = tid % 16.  RSTEP (a multiple
        // of 8) rows further is RSTEP / 8 swizzle atoms (1024 B each) further with the same XOR pattern: cells adst0 + 128 RSTEP i.
        constexpr int RSTEP = PROD_WARPS * 2, APT = R / RSTEP;
        const int arow0 = tid >> 4, ag = tid & 15;
        const bool acol_ok = ci0 + ag * 8 < p.cin;
        const uint32_t adst0 = (uint32_t)(ag >> 3) * SLAB + (uint32_t)arow0 * 128u + (uint32_t)(((ag & 7) ^ (arow0 & 7)) << 4);
        const uint32_t x_row_bytes = (uint32_t)p.cin * 2u, dy_row_bytes = (uint32_t)p.cout * 2u;
        const uint8_t *acol_hi = reinterpret_cast<const uint8_t *>(p.x_hi) + (acol_ok ? (ci0 + ag * 8) * 2 : 0);
        const uint8_t *acol_lo = reinterpret_cast<const uint8_t *>(p.x_lo) + (acol_ok ? (ci0 + ag * 8) * 2 : 0);
        const int brow0 = tid / GB, bg0 = tid - brow0 * GB;                                      // cell tid of a dY row block
        const int brow_step = (PROD_WARPS * 32) / GB, bg_step = PROD_WARPS * 32 - brow_step * GB;  // ... and the step to the next cell
        // Neighbour rows of every offset of the group, one row block AHEAD (register double buffer): the L2 round trip of the index
        // loads is covered by the copies of the current block instead of stalling all producers once per block.
        int nxt[MAX_KG][APT];
        auto load_indices = [&](int r0) {
#pragma unroll
            for (int kq = 0; kq < MAX_KG; ++kq) {
#pragma unroll
                for (int i = 0; i < APT; ++i) {
                    const int j = r0 + arow0 + RSTEP * i;
                    nxt[kq][i] = (kq < nk && j < r_end) ? __ldg(p.nbr + (size_t)(k0 + kq) * p.n_rows + j) : -1;
                }
            }
        };
        load_indices(r_begin);
        long long pw_b = 0, pw_a = 0;
        const long long pt0 = PROF ? clock64() : 0;
        for (int r0 = r_begin; r0 < r_end; r0 += R) {
            const uint32_t act = active(r0);
            int src[MAX_KG][APT];
#pragma unroll
            for (int kq = 0; kq < MAX_KG; ++kq)
#pragma unroll
                for (int i = 0; i < APT; ++i) src[kq][i] = nxt[kq][i];
            if (r0 + R < r_end) load_indices(r0 + R);
            if (!act) continue;
            touched |= act;
            // ---- B item: 64 rows of dY.  Cell e = tid + 128 n of the row block is 16-byte chunk g = e % GB of row e / GB; without a row
            // permutation the block is one contiguous range of dY, so the source of a cell is base + 16 e and only the swizzled
            // destination needs (row, g) — stepped without a division.
            {
                const long long tw = PROF ? clock64() : 0;
                mbar_wait(smem_u32(&b_empty[bs]), bpar ^ 1, 0);
                if (PROF) pw_b += clock64() - tw;
            }
            {
                const uint32_t slot = b_base + (uint32_t)bs * b_slot_bytes;
                const int rows_here = min(R, r_end - r0);
                if (p.dy_rows == nullptr) {
                    const uint8_t *src_hi = reinterpret_cast<const uint8_t *>(p.dy_hi) + (size_t)r0 * dy_row_bytes;
                    const uint8_t *src_lo = reinterpret_cast<const uint8_t *>(p.dy_lo) + (size_t)r0 * dy_row_bytes;
                    int row = brow0, g = bg0;
                    for (int e = tid; e < R * GB; e += PROD_WARPS * 32) {
                        const bool ign = row >= rows_here;
                        const uint32_t dst = slot + (uint32_t)(g >> 3) * SLAB + (uint32_t)row * 128u + (uint32_t)(((g & 7) ^ (row & 7)) << 4);
                        const uint32_t off = ign ? 0u : (uint32_t)e * 16u;
                        cp_async16_pred(dst, src_hi + off, ign);
                        if (PASSES == 3) cp_async16_pred(dst + b_plane, src_lo + off, ign);
                        row += brow_step;
                        g += bg_step;
                        if (g >= GB) {
                            g -= GB;
                            ++row;
                        }
                    }
                } else {
                    for (int it = tid; it < R * GB; it += PROD_WARPS * 32) {
                        const int row = it / GB, g = it - row * GB;
                        const bool ign = row >= rows_here;
                        const uint32_t dst = slot + (uint32_t)(g >> 3) * SLAB + (uint32_t)row * 128u + (uint32_t)(((g & 7) ^ (row & 7)) << 4);
                        const uint32_t jr = ign ? 0u : (uint32_t)__ldg(p.dy_rows + r0 + row);
                        const uint64_t off = (uint64_t)jr * dy_row_bytes + (uint32_t)g * 16u;
                        cp_async16_pred(dst, reinterpret_cast<const uint8_t *>(p.dy_hi) + off, ign);
                        if (PASSES == 3) cp_async16_pred(dst + b_plane, reinterpret_cast<const uint8_t *>(p.dy_lo) + off, ign);
                    }
                }
                cp_async_arrive_noinc(smem_u32(&b_full[bs]));
                if (++bs == p.b_slots) {
                    bs = 0;
                    bpar ^= 1u;
                }
            }
            // ---- A items: gathered X rows for every active offset of the group, one plane per slot
#pragma unroll
            for (int kq = 0; kq < MAX_KG; ++kq) {
                if (kq >= nk || !((act >> (k0 + kq)) & 1u)) continue;
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) {
                    const uint8_t *plane = pl == 0 ? acol_hi : acol_lo;
                    const int m = kq % NMW;  // compile-time after unrolling
                    const long long tw = PROF ? clock64() : 0;
                    mbar_wait(smem_u32(&a_empty[m][as[m]]), apar[m] ^ 1, 1);
                    if (PROF) pw_a += clock64() - tw;
                    const uint32_t slot = a_base + (uint32_t)(m * p.a_slots + as[m]) * A_SLOT + adst0;
#pragma unroll
                    for (int i = 0; i < APT; ++i) {
                        const bool ign = src[kq][i] < 0 || !acol_ok;
                        const uint64_t off = (uint64_t)(uint32_t)max(src[kq][i], 0) * x_row_bytes;
                        cp_async16_pred(slot + (uint32_t)i * (128u * RSTEP), plane + off, ign);
                    }
                    cp_async_arrive_noinc(smem_u32(&a_full[m][as[m]]));
                    if (++as[m] == p.a_slots) {
                        as[m] = 0;
                        apar[m] ^= 1u;
                    }
                }
            }
        }
        cp_async_wait_all();
        const long long pt1 = PROF ? clock64() : 0;

        // ------------------------------------------------------------------ epilogue: lanes = ci, columns = co
        if (touched) {
            mbar_wait(smem_u32(&acc_full), 0, 2);
            tc_fence_after();
            // warp w may touch TMEM lanes 32 (w % 4) .. + 31; the PROD_WARPS / 4 warps of a lane quarter take alternate 16-column chunks
            const int ci = (warp & 3) * 32 + lane;
            const bool ci_ok = ci0 + ci < p.cin;
            constexpr int CW = PROD_WARPS / 4;
            for (int kq = 0; kq < nk; ++kq) {
                if (!((touched >> (k0 + kq)) & 1u)) continue;
                float *drow = p.dw + ((size_t)(k0 + kq) * p.cin + ci0 + ci) * p.cout;
                const uint32_t taddr = tmem_base + (uint32_t)(kq * p.npad) + ((uint32_t)((warp & 3) * 32) << 16);
                for (int col = (warp >> 2) * 16; col < p.cout; col += 16 * CW) {
                    float acc[16];
                    tmem_ld16(taddr + (uint32_t)col, acc);
                    if (ci_ok) {
#pragma unroll
                        for (int e = 0; e < 16; e += 4) atomicAdd(reinterpret_cast<float4 *>(drow + col + e), make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]));
                    }
                }
            }
        }
        if (PROF && p.prof != nullptr && tid == 0) {
            long long *o = p.prof + (size_t)blockIdx.x * 16;
            o[0] = pt1 - pt0;          // producer loop
            o[1] = pw_b;               // ... waiting for a free dY slot
            o[2] = pw_a;               // ... waiting for a free X slot
            o[3] = clock64() - pt1;    // epilogue (incl. the wait for the last MMAs)
        }
    } else {
        // ------------------------------------------------------------------ MMA issuers (uniform warps, one elected lane issues)
        // tcgen05.mma issue blocks for about the instruction's execution time, so a single issuing warp leaves the tensor pipe idle
        // while it waits for a slot, fences and books (measured: 58 % issue, 42 % other).  Two warps, each with its own offsets,
        // accumulators and A ring, keep the pipe fed; both consume every dY tile.
        const int mw = warp - MMA_WARP;
        const uint32_t idesc = idesc_bf16(p.nmma, true, true);  // N = cout rounded up to 16: columns past it are never read back
        const uint64_t a_desc0 = desc_mn_sw128(a_base + (uint32_t)(mw * p.a_slots) * A_SLOT, SLAB), b_desc0 = desc_mn_sw128(b_base, SLAB);
        int as = 0, bs = 0;
        uint32_t apar = 0, bpar = 0, mine = 0;  // mine: offsets of this warp that received an MMA
        long long mw_b = 0, mw_a = 0, m_issue = 0, n_slots = 0;
        const long long mt0 = PROF ? clock64() : 0;
        for (int r0 = r_begin; r0 < r_end; r0 += R) {
            const uint32_t act = active(r0);
            if (!act) continue;
            touched |= act;
            const long long twb = PROF ? clock64() : 0;
            mbar_wait(smem_u32(&b_full[bs]), bpar, 3);
            if (PROF) mw_b += clock64() - twb;
            const uint64_t db_hi = b_desc0 + (uint64_t)((uint32_t)(bs * b_slot_bytes) >> 4);
            const uint64_t db_lo = db_hi + (uint64_t)((uint32_t)b_plane >> 4);
            for (int kq = mw; kq < nk; kq += NMW) {
                const int k = k0 + kq;
                if (!((act >> k) & 1u)) continue;
                const uint32_t acc = tmem_base + (uint32_t)(kq * p.npad);
                const uint32_t first = (mine >> k) & 1u;
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) {
                    const long long twa = PROF ? clock64() : 0;
                    mbar_wait(smem_u32(&a_full[mw][as]), apar, 4);
                    const long long ti0 = PROF ? clock64() : 0;
                    if (PROF) {
                        mw_a += ti0 - twa;
                        ++n_slots;
                    }
                    fence_proxy_async();  // cp.async writes of the producers (generic proxy) -> tcgen05 reads (async proxy)
                    tc_fence_after();
                    const uint64_t da = a_desc0 + (uint64_t)((uint32_t)(as * A_SLOT) >> 4);
                    if (elect_one()) {
                        if (pl == 0) {
#pragma unroll
                            for (int kk = 0; kk < R / 16; ++kk) {
                                const uint64_t adv = (uint64_t)(kk * (2048 >> 4));  // two 8-row groups
                                umma(acc, da + adv, db_hi + adv, idesc, first | (kk != 0));
                                if (PASSES == 3) umma(acc, da + adv, db_lo + adv, idesc, 1);
                            }
                        } else {
#pragma unroll
                            for (int kk = 0; kk < R / 16; ++kk) {
                                const uint64_t adv = (uint64_t)(kk * (2048 >> 4));
                                umma(acc, da + adv, db_hi + adv, idesc, 1);
                            }
                        }
                        umma_commit(smem_u32(&a_empty[mw][as]));
                    }
                    __syncwarp();
                    if (PROF) m_issue += clock64() - ti0;
                    if (++as == p.a_slots) {
                        as = 0;
                        apar ^= 1u;
                    }
                }
                mine |= 1u << k;
            }
            if (elect_one()) umma_commit(smem_u32(&b_empty[bs]));  // arrives once this warp's MMAs on the tile are done (at once if none)
            __syncwarp();
            if (++bs == p.b_slots) {
                bs = 0;
                bpar ^= 1u;
            }
        }
        if (touched && elect_one()) umma_commit(smem_u32(&acc_full));
        __syncwarp();
        if (PROF && p.prof != nullptr && lane == 0 && mw == 0) {
            long long *o = p.prof + (size_t)blockIdx.x * 16;
            o[4] = clock64() - mt0;  // MMA warp loop
            o[5] = mw_b;             // ... waiting for dY
            o[6] = mw_a;             // ... waiting for gathered X
            o[7] = m_issue;          // ... fence + issue + commit
            o[8] = n_slots;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, (uint32_t)p.acc_cols);
}


// out[j, :] = in[order[j], :] for both bf16 planes (rows of c channels, c % 8 == 0): dY brought into the row order of a
// pattern-ordered table, so that the weight-gradient kernel streams its dY tiles instead of gathering scattered rows.
__global__ void __launch_bounds__(256)
k_permute_planes(const uint4 *__restrict__ hi, const uint4 *__restrict__ lo, const int32_t *__restrict__ order, int n, int chunks,
                 uint4 *__restrict__ hi_out, uint4 *__restrict__ lo_out) {
    const long long total = (long long)n * chunks;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e / chunks), g = (int)(e - (long long)j * chunks);
        const size_t src = (size_t)__ldg(order + j) * chunks + g;
        hi_out[e] = hi[src];
        if (lo != nullptr) lo_out[e] = lo[src];
    }
}

}  // namespace wg
}  // namespace us3d

using namespace us3d;

static long long *g_wgrad_prof = nullptr;

extern "C" {

/* debug hook (include/us3d_debug.h): per-CTA cycle counters of the weight-gradient kernel's roles, 16 int64 per CTA */
void us3d_debug_set_prof_wgrad(void *buf) { g_wgrad_prof = (long long *)buf; }

int us3d_spconv_wgrad_planes(const void *x_hi, const void *x_lo, const void *dy_hi, const void *dy_lo, const int32_t *nbr,
                             int n_rows, int kvol, float *dw, int cin, int cout, int passes, const uint32_t *tile_mask,
                             const int32_t *dy_rows, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "spconv_wgrad_planes: kvol %d out of range", kvol);
    US3D_CHECK_ARG(passes == 1 || passes == 3, "spconv_wgrad_planes: passes must be 1 or 3");
    US3D_CHECK_ARG(us3d_spconv_wgrad_tc_supported(cin, cout), "spconv_wgrad_planes: unsupported channel counts %d -> %d", cin, cout);
    US3D_CHECK_ARG(x_hi && dy_hi && (passes == 1 || (x_lo && dy_lo)), "spconv_wgrad_planes: missing plane");
    if (n_rows == 0) return 0;
    wg::Params p;
    p.x_hi = (const __nv_bfloat16 *)x_hi; p.x_lo = (const __nv_bfloat16 *)x_lo;
    p.dy_hi = (const __nv_bfloat16 *)dy_hi; p.dy_lo = (const __nv_bfloat16 *)dy_lo;
    p.nbr = nbr; p.n_rows = n_rows; p.kvol = kvol; p.dw = dw; p.cin = cin; p.cout = cout; p.tile_mask = tile_mask; p.dy_rows = dy_rows;
    p.npad = ceil_div(cout, 64) * 64;
    p.nmma = ceil_div(cout, 16) * 16;
    p.mblks = ceil_div(cin, 128);
    int kg = 512 / p.npad;
    if (kg > wg::MAX_KG) kg = wg::MAX_KG;
    if (kg > kvol) kg = kvol;
    p.kg = kg;
    p.ngroups = ceil_div(kvol, kg);
    int acc = 32;
    while (acc < kg * p.npad) acc <<= 1;
    p.acc_cols = acc;
    // one CTA per SM, whole waves
    int units = p.ngroups * p.mblks;
    int splits = num_sms() / units;
    if (splits < 1) splits = 1;
    int max_splits = ceil_div(n_rows, 512);
    if (splits > max_splits) splits = max_splits;
    p.rows_per_split = ceil_div(ceil_div(n_rows, splits), 128) * 128;
    p.splits = ceil_div(n_rows, p.rows_per_split);
    const int npl = passes == 3 ? 2 : 1;
    const int a_slot = 2 * wg::SLAB, b_slot = npl * (p.npad / 64) * wg::SLAB;
    const int budget = 220 * 1024;
    p.b_slots = 2;
    p.a_slots = (budget - p.b_slots * b_slot) / a_slot / wg::NMW;  // per A ring (one ring per MMA-issuing warp)
    if (p.a_slots > wg::MAX_A) p.a_slots = wg::MAX_A;
    US3D_CHECK_ARG(p.a_slots >= 2, "spconv_wgrad_planes: operand slots do not fit in shared memory (cout %d)", cout);
    const size_t smem = (size_t)(wg::NMW * p.a_slots) * a_slot + (size_t)p.b_slots * b_slot + 1024;
    static bool attr_done[64] = {};  // the opt-in for > 48 KB of dynamic shared memory is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        US3D_CUDA(cudaFuncSetAttribute(wg::k_wgrad<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        US3D_CUDA(cudaFuncSetAttribute(wg::k_wgrad<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        US3D_CUDA(cudaFuncSetAttribute(wg::k_wgrad<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        attr_done[dev] = true;
    }
    const int grid = p.ngroups * p.mblks * p.splits;
    {
        ProfScope prof(st, 1, n_rows, n_rows, kvol, cin, cout);
        p.prof = g_wgrad_prof;
        if (passes == 3 && p.prof != nullptr)
            US3D_CUDA(launch_pdl(wg::k_wgrad<3, true>, dim3(grid), dim3(wg::THREADS), smem, st, p));
        else if (passes == 3)
            US3D_CUDA(launch_pdl(wg::k_wgrad<3, false>, dim3(grid), dim3(wg::THREADS), smem, st, p));
        else
            US3D_CUDA(launch_pdl(wg::k_wgrad<1, false>, dim3(grid), dim3(wg::THREADS), smem, st, p));
    }
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_permute_planes(const void *hi, const void *lo, const int32_t *order, int n, int c, void *hi_out, void *lo_out, void *stream_) {
    US3D_CHECK_ARG(n >= 0 && c > 0 && c % 8 == 0, "permute_planes: channels must be a multiple of 8");
    US3D_CHECK_ARG(hi && hi_out && order && ((lo == nullptr) == (lo_out == nullptr)), "permute_planes: missing plane");
    if (n == 0) return 0;
    const int chunks = c / 8;
    long long blocks = ((long long)n * chunks + 255) / 256;
    if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
    wg::k_permute_planes<<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>((const uint4 *)hi, (const uint4 *)lo, order, n, chunks,
                                                                        (uint4 *)hi_out, (uint4 *)lo_out);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
