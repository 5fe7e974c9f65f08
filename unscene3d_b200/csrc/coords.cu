// Coordinate maps and kernel maps (SURVEY §8(a) A1–A3): voxel hashing, stride/unique, neighbour tables.
//
// The reference gets these from MinkowskiEngine's CoordinateManager (implicit inside every
// MinkowskiConvolution / pooling forward; requested via ME.KernelGenerator,
// /root/reference/models/modules/common.py:137-144).  Here they are integer kernels over an
// open-addressing hash table that lives in global memory — on B200 the whole table (<= 16 B/slot,
// 2 slots/voxel) stays L2-resident, so a probe costs an L2 hit, not an HBM access.
//
// Canonical order of a de-duplicated map = order of first occurrence among the input rows; that is
// obtained without sorting: every row atomicMin's its index into the slot of its key, winners are
// flagged, an exclusive scan of the flags is the output row.
#include <limits.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include <stdlib.h>
#include "common.cuh"

namespace us3d {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- per-launch profiling -----------------------------------------------------------------------
bool g_profile = false;
bool g_pdl = [] {
    const char *e = getenv("US3D_PDL");
    return !(e && e[0] == '0');
}();
namespace {
struct ProfRec {
    cudaEvent_t s, e;
    int meta[7];
};
std::vector<ProfRec> g_prof_recs;
std::mutex g_prof_mu;
int g_prof_tag = 0;
}  // namespace

void prof_begin(cudaStream_t st, int kind, int n_in, int n_rows, int kvol, int cin, int cout) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    cudaEventCreate(&r.s);
    cudaEventCreate(&r.e);
    r.meta[0] = kind == 0 ? g_prof_tag : kind;
    r.meta[1] = n_in; r.meta[2] = n_rows; r.meta[3] = kvol; r.meta[4] = cin; r.meta[5] = cout; r.meta[6] = 0;
    cudaEventRecord(r.s, st);
    g_prof_recs.push_back(r);
}

void prof_end(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().e, st);
}

constexpr int kScanThreads = 512;
constexpr int kScanItems = 4;
constexpr int kScanChunk = kScanThreads * kScanItems;  // rows per scan block
constexpr int kMaxScanBlocks = 4096;

__global__ void k_hash_clear(uint64_t *keys, int32_t *vals, int cap) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) {
        keys[i] = kEmptyKey;
        vals[i] = INT_MAX;
    }
}

// one thread per input row: claim (or find) the slot of the row's key, remember it, vote for "first".
__global__ void k_insert(const int4 *__restrict__ coords, int n, int tsx, int tsy, int tsz, uint64_t *keys,
                         int32_t *vals, uint32_t mask, int32_t *__restrict__ slot_of_row, int *__restrict__ bad) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    int4 c = coords[row];
    int x = floor_to(c.y, tsx), y = floor_to(c.z, tsy), z = floor_to(c.w, tsz);
    if (!key_in_range(c.x, x, y, z)) {
        atomicExch(bad, 1);
        slot_of_row[row] = -1;
        return;
    }
    uint64_t key = pack_key(c.x, x, y, z);
    uint32_t slot = hash_key(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS((unsigned long long *)&keys[slot], (unsigned long long)kEmptyKey,
                                            (unsigned long long)key);
        if (prev == kEmptyKey || prev == key) break;
        slot = (slot + 1) & mask;
    }
    atomicMin(&vals[slot], row);
    slot_of_row[row] = (int32_t)slot;
}

__global__ void k_flag_first(int n, const int32_t *__restrict__ slot_of_row, const int32_t *__restrict__ vals,
                             int32_t *__restrict__ flag) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    int s = slot_of_row[row];
    flag[row] = (s >= 0 && vals[s] == row) ? 1 : 0;
}

// block-local exclusive scan (in place) + block total
__global__ void __launch_bounds__(kScanThreads) k_scan_blocks(int32_t *data, int n, int32_t *block_sums) {
    __shared__ int32_t warp_tot[kScanThreads / 32];
    int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
    int v[kScanItems], t = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? data[base + i] : 0;
        t += v[i];
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kScanThreads / 32 ? warp_tot[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += o;
        }
        if (lane < kScanThreads / 32) warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    int excl = inc - t + (warp ? warp_tot[warp - 1] : 0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) data[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == kScanThreads - 1) block_sums[blockIdx.x] = warp_tot[kScanThreads / 32 - 1];
}

// single block: exclusive scan of block sums in place, grand total to sums[nb]
__global__ void __launch_bounds__(1024) k_scan_sums(int32_t *sums, int nb) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < nb; start += 1024) {
        int i = start + threadIdx.x;
        int v = i < nb ? sums[i] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int o = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += o;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        int carry = carry_s;
        int excl = carry + inc - v + (warp ? warp_tot[warp - 1] : 0);
        if (i < nb) sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nb] = carry_s;
}

// winners write their unique row; the table value becomes the unique row index
__global__ void k_emit_unique(const int4 *__restrict__ coords, int n, int tsx, int tsy, int tsz,
                              const int32_t *__restrict__ slot_of_row, const int32_t *__restrict__ scan,
                              const int32_t *__restrict__ block_sums, int32_t *vals, int4 *__restrict__ out_coords,
                              int32_t *__restrict__ out_first) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    int s = slot_of_row[row];
    if (s < 0 || vals[s] != row) return;  // only winners still see their own row index here
    int r = scan[row] + block_sums[row / kScanChunk];
    int4 c = coords[row];
    out_coords[r] = make_int4(c.x, floor_to(c.y, tsx), floor_to(c.z, tsy), floor_to(c.w, tsz));
    out_first[r] = row;
    // no other thread reads vals[s] in this kernel except through the `!= row` test above, and a
    // unique-row index r <= row, r == row only for an untouched prefix, so the test stays correct.
    vals[s] = r | 0x40000000;  // tagged so that a concurrent loser can never mistake it for its row
}

__global__ void k_untag_inverse(int n, const int32_t *__restrict__ slot_of_row, int32_t *vals,
                                int32_t *__restrict__ inverse, int phase) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    int s = slot_of_row[row];
    if (phase == 0) {
        if (s >= 0 && (vals[s] & 0x40000000)) {
            // every row of the key clears the tag with the same value: benign
            vals[s] = vals[s] & 0x3FFFFFFF;
        }
    } else {
        inverse[row] = s >= 0 ? vals[s] : -1;
    }
}

struct Offsets {
    int32_t v[US3D_MAX_KVOL * 3];
};

// grid.y = kernel offset, one thread per query row
__global__ void k_kernel_map(const int4 *__restrict__ query, int n_q, Offsets offs, const uint64_t *__restrict__ keys,
                             const int32_t *__restrict__ vals, uint32_t mask, int32_t *__restrict__ nbr,
                             uint32_t *tile_mask, int tile_rows) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    int hit = -1;
    if (q < n_q) {
        int4 c = query[q];
        int x = c.y + offs.v[k * 3 + 0], y = c.z + offs.v[k * 3 + 1], z = c.w + offs.v[k * 3 + 2];
        if (key_in_range(c.x, x, y, z)) hit = hash_lookup(keys, vals, mask, pack_key(c.x, x, y, z));
        nbr[(size_t)k * n_q + q] = hit;
    }
    if (tile_mask != nullptr) {
        unsigned any = __ballot_sync(0xffffffffu, hit >= 0);
        if (any && (threadIdx.x & 31) == 0 && q < n_q) atomicOr(&tile_mask[q / tile_rows], 1u << k);
    }
}


// ---- rows ordered by neighbour pattern ------------------------------------------------------------
// The tcgen05 kernels work on tiles of 128 output rows and skip a kernel offset only when NO row of the tile has a
// neighbour there.  On voxelised surfaces every offset is present somewhere in any 128 spatially consecutive rows
// (measured on the 200k-voxel scene: 100 % of (tile, offset) pairs active at a pair density of 49 %), so the rows of a
// large map are instead grouped by their presence pattern (bit k = neighbour at offset k): sorted by the pattern with
// the rarest offsets as the most significant bits, a tile's union pattern covers 67 % of the offsets.
// Pass 1: pattern per row + how often each offset is present.
__global__ void __launch_bounds__(256) k_pattern_count(const int32_t *__restrict__ nbr, int n, int kvol, uint32_t *__restrict__ pattern,
                                                       unsigned *counts) {
    __shared__ unsigned s_cnt[32];
    if (threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int j0 = blockIdx.x * blockDim.x; j0 < n; j0 += gridDim.x * blockDim.x) {
        const int j = j0 + threadIdx.x;
        uint32_t pat = 0;
        for (int k = 0; k < kvol; ++k) {
            bool hit = j < n && nbr[(size_t)k * n + j] >= 0;
            pat |= (uint32_t)hit << k;
            unsigned b = __ballot_sync(0xffffffffu, hit);
            if (b && (threadIdx.x & 31) == 0) atomicAdd(&s_cnt[k], __popc(b));  // block-local: 27 global atomics per block
        }
        if (j < n) pattern[j] = pat;
    }
    __syncthreads();
    if (threadIdx.x < kvol && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]);
}
// Pass 2: sort key = pattern with its bits permuted by presence count (most frequent offset -> bit 0; ties by offset index)
__global__ void k_pattern_key(const uint32_t *__restrict__ pattern, int n, int kvol, const unsigned *__restrict__ counts,
                              int32_t *__restrict__ key) {
    __shared__ int pos[32];
    if (threadIdx.x < 32) {
        int k = threadIdx.x, r = 0;
        if (k < kvol) {
            unsigned ck = counts[k];
            for (int q = 0; q < kvol; ++q) {
                unsigned cq = counts[q];
                r += (cq > ck) || (cq == ck && q < k);
            }
        }
        pos[k] = r;
    }
    __syncthreads();
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t pat = pattern[j], out = 0;
    for (int k = 0; k < kvol; ++k) out |= ((pat >> k) & 1u) << pos[k];
    key[j] = (int32_t)out;
}
// nbr_out[k, j] = nbr[k, order[j]] + the tile masks of the re-ordered table
__global__ void k_reorder_table(const int32_t *__restrict__ nbr, int n, const int32_t *__restrict__ order, int32_t *__restrict__ nbr_out,
                                uint32_t *tile_mask, int tile_rows) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    int hit = -1;
    if (j < n) {
        hit = nbr[(size_t)k * n + order[j]];
        nbr_out[(size_t)k * n + j] = hit;
    }
    unsigned any = __ballot_sync(0xffffffffu, hit >= 0);
    if (any && (threadIdx.x & 31) == 0 && j < n) atomicOr(&tile_mask[j / tile_rows], 1u << k);
}

}  // namespace us3d

using namespace us3d;

extern "C" {

int us3d_abi_version(void) { return US3D_ABI_VERSION; }
const char *us3d_last_error(void) { return g_err; }
long long us3d_launch_count(void) { return g_launches.load(); }

/* Profiling hooks (debug surface, used by bench.py): start collecting, tag the next gather launches (0 forward,
 * 2 input gradient), stop = synchronise and return up to `cap` records as meta[7 * i .. ] = (kind, n_in, n_rows, kvol,
 * cin, cout, 0) and ms[i]; kind 1 = weight gradient. */
/* programmatic dependent launch of the conv / BatchNorm kernels on (default, or US3D_PDL != 0) / off */
void us3d_debug_set_pdl(int on) { g_pdl = on != 0; }

void us3d_debug_profile_start(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &r : g_prof_recs) {
        cudaEventDestroy(r.s);
        cudaEventDestroy(r.e);
    }
    g_prof_recs.clear();
    g_profile = true;
}
void us3d_debug_profile_tag(int kind) { g_prof_tag = kind; }
int us3d_debug_profile_stop(int *meta, float *ms, int cap) {
    g_profile = false;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int n = 0;
    for (auto &r : g_prof_recs) {
        if (n < cap) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, r.s, r.e) != cudaSuccess) t = -1.f;
            for (int i = 0; i < 7; ++i) meta[7 * n + i] = r.meta[i];
            ms[n] = t;
            ++n;
        }
        cudaEventDestroy(r.s);
        cudaEventDestroy(r.e);
    }
    g_prof_recs.clear();
    cudaGetLastError();
    return n;
}
void us3d_reset_launch_count(void) { g_launches.store(0); }

int us3d_hash_capacity(int n) {
    long long need = 2LL * (n < 1 ? 1 : n);
    long long cap = 1024;
    while (cap < need) cap <<= 1;
    return cap > INT_MAX ? -1 : (int)cap;
}

int us3d_coords_unique(const int32_t *coords, int n, int tsx, int tsy, int tsz, uint64_t *keys, int32_t *vals, int cap,
                       int32_t *out_coords, int32_t *out_first, int32_t *inverse, int32_t *scratch, int *out_count_h,
                       void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n >= 0 && tsx > 0 && tsy > 0 && tsz > 0, "coords_unique: bad n/stride");
    US3D_CHECK_ARG(cap >= 2 * n && (cap & (cap - 1)) == 0, "coords_unique: capacity %d must be a power of two >= 2n", cap);
    int nb = ceil_div(n, kScanChunk);
    US3D_CHECK_ARG(nb <= kMaxScanBlocks - 2, "coords_unique: n=%d exceeds %d rows", n, (kMaxScanBlocks - 2) * kScanChunk);
    const int T = 256;
    k_hash_clear<<<ceil_div(cap, T), T, 0, st>>>(keys, vals, cap);
    US3D_LAUNCH_CHECK();
    if (n == 0) {
        *out_count_h = 0;
        US3D_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    int32_t *slot_of_row = scratch, *scan = scratch + n, *sums = scratch + 2 * n;  // sums[nb+1], bad flag at sums[nb+1]
    int *bad = sums + nb + 1;
    US3D_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
    uint32_t mask = (uint32_t)cap - 1;
    const int4 *c4 = reinterpret_cast<const int4 *>(coords);
    k_insert<<<ceil_div(n, T), T, 0, st>>>(c4, n, tsx, tsy, tsz, keys, vals, mask, slot_of_row, bad);
    US3D_LAUNCH_CHECK();
    k_flag_first<<<ceil_div(n, T), T, 0, st>>>(n, slot_of_row, vals, scan);
    US3D_LAUNCH_CHECK();
    k_scan_blocks<<<nb, kScanThreads, 0, st>>>(scan, n, sums);
    US3D_LAUNCH_CHECK();
    k_scan_sums<<<1, 1024, 0, st>>>(sums, nb);
    US3D_LAUNCH_CHECK();
    k_emit_unique<<<ceil_div(n, T), T, 0, st>>>(c4, n, tsx, tsy, tsz, slot_of_row, scan, sums, vals,
                                                reinterpret_cast<int4 *>(out_coords), out_first);
    US3D_LAUNCH_CHECK();
    k_untag_inverse<<<ceil_div(n, T), T, 0, st>>>(n, slot_of_row, vals, inverse, 0);
    US3D_LAUNCH_CHECK();
    k_untag_inverse<<<ceil_div(n, T), T, 0, st>>>(n, slot_of_row, vals, inverse, 1);
    US3D_LAUNCH_CHECK();
    int host[2] = {0, 0};
    US3D_CUDA(cudaMemcpyAsync(host, sums + nb, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    US3D_CUDA(cudaStreamSynchronize(st));
    US3D_CHECK_ARG(host[1] == 0, "coords_unique: coordinate outside the 10/18/18/18-bit key range");
    *out_count_h = host[0];
    return 0;
}

int us3d_kernel_map(const int32_t *query, int n_q, const int32_t *offsets_h, int kvol, const uint64_t *keys,
                    const int32_t *vals, int cap, int32_t *nbr, uint32_t *tile_mask, int tile_rows, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "kernel_map: kvol %d out of range", kvol);
    US3D_CHECK_ARG((cap & (cap - 1)) == 0, "kernel_map: capacity must be a power of two");
    US3D_CHECK_ARG(tile_mask == nullptr || (tile_rows > 0 && tile_rows % 32 == 0), "kernel_map: tile_rows must be a multiple of 32");
    if (n_q == 0) return 0;
    Offsets offs;
    memset(&offs, 0, sizeof(offs));
    memcpy(offs.v, offsets_h, sizeof(int32_t) * 3 * kvol);
    if (tile_mask) US3D_CUDA(cudaMemsetAsync(tile_mask, 0, sizeof(uint32_t) * ceil_div(n_q, tile_rows), st));
    const int T = 256;
    dim3 grid(ceil_div(n_q, T), kvol);
    k_kernel_map<<<grid, T, 0, st>>>(reinterpret_cast<const int4 *>(query), n_q, offs, keys, vals, (uint32_t)cap - 1, nbr,
                                     tile_mask, tile_rows);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_neighbour_pattern_keys(const int32_t *nbr, int n_rows, int kvol, uint32_t *scratch, int32_t *keys, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "neighbour_pattern_keys: kvol %d out of range", kvol);
    if (n_rows == 0) return 0;
    unsigned *counts = scratch + n_rows;  // scratch: uint32[n_rows + 32] (patterns, then the per-offset counts)
    US3D_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned) * 32, st));
    int blocks = ceil_div(n_rows, 256);
    if (blocks > num_sms() * 4) blocks = num_sms() * 4;
    k_pattern_count<<<blocks, 256, 0, st>>>(nbr, n_rows, kvol, scratch, counts);
    US3D_LAUNCH_CHECK();
    k_pattern_key<<<ceil_div(n_rows, 256), 256, 0, st>>>(scratch, n_rows, kvol, counts, keys);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_kernel_map_reorder(const int32_t *nbr, int n_rows, int kvol, const int32_t *order, int32_t *nbr_out, uint32_t *tile_mask,
                            int tile_rows, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(kvol >= 1 && kvol <= US3D_MAX_KVOL, "kernel_map_reorder: kvol %d out of range", kvol);
    US3D_CHECK_ARG(tile_mask != nullptr && tile_rows > 0 && tile_rows % 32 == 0, "kernel_map_reorder: tile_rows must be a multiple of 32");
    if (n_rows == 0) return 0;
    US3D_CUDA(cudaMemsetAsync(tile_mask, 0, sizeof(uint32_t) * ceil_div(n_rows, tile_rows), st));
    k_reorder_table<<<dim3(ceil_div(n_rows, 256), kvol), 256, 0, st>>>(nbr, n_rows, order, nbr_out, tile_mask, tile_rows);
    US3D_LAUNCH_CHECK();
    return 0;
}

/* HOST function (host pointers, no CUDA context: fork-safe, as ME.utils.sparse_quantize in the reference's DataLoader workers,
 * datasets/utils.py:266-270, 403-408).  Unique rows of coords_h [n, d] (int32, d <= 8) in order of first occurrence:
 * first_h[u] = input row of unique row u (ascending), inverse_h[i] = unique row of input row i.  Returns the number of unique
 * rows, or a negative error. */
int us3d_coords_unique_h(const int32_t *coords_h, int n, int d, int32_t *first_h, int32_t *inverse_h) {
    US3D_CHECK_ARG(n >= 0 && d >= 1 && d <= 8, "coords_unique_h: bad shape");
    if (n == 0) return 0;
    size_t cap = 1024;
    while (cap < (size_t)n * 2) cap <<= 1;
    std::vector<int32_t> slot(cap, -1);  // unique row stored in each slot
    int count = 0;
    for (int i = 0; i < n; ++i) {
        const int32_t *c = coords_h + (size_t)i * d;
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int a = 0; a < d; ++a) {
            h ^= (uint64_t)(uint32_t)c[a] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
            h *= 0xBF58476D1CE4E5B9ull;
        }
        size_t p = (size_t)(h ^ (h >> 31)) & (cap - 1);
        for (;;) {
            const int32_t u = slot[p];
            if (u < 0) {
                slot[p] = count;
                first_h[count] = i;
                inverse_h[i] = count++;
                break;
            }
            if (memcmp(coords_h + (size_t)first_h[u] * d, c, sizeof(int32_t) * d) == 0) {
                inverse_h[i] = u;
                break;
            }
            p = (p + 1) & (cap - 1);
        }
    }
    return count;
}

}  // extern "C"
