// Shared helpers for libus3d (sm_100a).  Host code of the C ABI lives next to each kernel file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/us3d.h"

namespace us3d {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define US3D_CHECK_ARG(cond, ...)              \
    do {                                       \
        if (!(cond)) {                         \
            ::us3d::set_error(__VA_ARGS__);    \
            return -1;                         \
        }                                      \
    } while (0)

#define US3D_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ::us3d::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)

#define US3D_LAUNCH_CHECK()                                                                    \
    do {                                                                                       \
        ::us3d::count_launch();                                                                \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) {                                                              \
            ::us3d::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return -3;                                                                         \
        }                                                                                      \
    } while (0)

inline int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// ---- per-launch profiling (bench.py roofline leg) ------------------------------------------------
// When enabled, the convolution entry points bracket their kernel launch with CUDA events recorded on the launch
// stream from inside the library — the start event sits a few microseconds of host time before the kernel, so a
// host-bound step does not leak its launch gaps into the kernel's duration.  kind: 0 forward / input gradient
// (as tagged by the caller), 1 weight gradient.
void prof_begin(cudaStream_t st, int kind, int n_in, int n_rows, int kvol, int cin, int cout);
void prof_end(cudaStream_t st);
extern bool g_profile;
struct ProfScope {
    cudaStream_t st;
    bool on;
    ProfScope(cudaStream_t s, int kind, int n_in, int n_rows, int kvol, int cin, int cout) : st(s), on(g_profile) {
        if (on) prof_begin(st, kind, n_in, n_rows, kvol, cin, cout);
    }
    ~ProfScope() {
        if (on) prof_end(st);
    }
};

// ---- programmatic dependent launch ---------------------------------------------------------------
// The step is ~560 short launches of this library on one stream.  A kernel launched through launch_pdl may be scheduled while
// its predecessor is still draining: it runs its prologue (barrier init, TMEM allocation, shared-memory clears) and then blocks in
// pdl_wait() until the predecessor has completed and its writes are visible; pdl_trigger() in the predecessor is what lets the
// scheduler start it early.  Every kernel launched this way executes pdl_wait() before its first global access.  US3D_PDL=0
// (or us3d_debug_set_pdl(0)) launches the same kernels without the attribute; the two instructions are then no-ops.
extern bool g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- voxel keys -------------------------------------------------------------------------------
// b:10 | x:18 | y:18 | z:18, each spatial field biased by 2^17 so negative coordinates order correctly.
constexpr int kAxisBits = 18;
constexpr int kAxisBias = 1 << (kAxisBits - 1);
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ __forceinline__ uint64_t pack_key(int b, int x, int y, int z) {
    return ((uint64_t)(uint32_t)b << (3 * kAxisBits)) | ((uint64_t)(uint32_t)(x + kAxisBias) << (2 * kAxisBits)) |
           ((uint64_t)(uint32_t)(y + kAxisBias) << kAxisBits) | (uint64_t)(uint32_t)(z + kAxisBias);
}

__host__ __device__ __forceinline__ bool key_in_range(int b, int x, int y, int z) {
    return b >= 0 && b < (1 << 10) && x >= -kAxisBias && x < kAxisBias && y >= -kAxisBias && y < kAxisBias &&
           z >= -kAxisBias && z < kAxisBias;
}

__device__ __forceinline__ uint32_t hash_key(uint64_t k) {  // murmur3 finaliser
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
}

// floor(a / s) * s for s > 0, also for negative a
__host__ __device__ __forceinline__ int floor_to(int a, int s) {
    int q = a / s, r = a % s;
    if (r < 0) --q;
    return q * s;
}

__device__ __forceinline__ int hash_lookup(const uint64_t *__restrict__ keys, const int32_t *__restrict__ vals,
                                           uint32_t mask, uint64_t key) {
    uint32_t slot = hash_key(key) & mask;
    while (true) {
        uint64_t k = keys[slot];
        if (k == key) return vals[slot];
        if (k == kEmptyKey) return -1;
        slot = (slot + 1) & mask;
    }
}

}  // namespace us3d
