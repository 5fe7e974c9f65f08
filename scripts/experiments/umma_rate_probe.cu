// How fast does one SM retire tcgen05.mma (kind::f16, bf16 -> fp32, cta_group::1, M = 128, K = 16) as a function of N, with
// both operands in shared memory (SS mode, K-major SWIZZLE_128B — the layout of the sparse-conv kernels)?  One CTA per SM,
// one elected thread issues `iters` MMAs back to back and commits; clock64 around issue + completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I unscene3d_b200/csrc scripts/experiments/umma_rate_probe.cu -o /tmp/umma_rate_probe -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace us3d::tcx;

// mode 0: all MMAs into one accumulator; mode 1: alternate two accumulators; mode 2: A operand advances through 4 K-steps of a
// 64-channel slot and B likewise (like the conv kernel); writers > 0: that many warps stream 16-byte st.shared into a scratch
// region meanwhile (stand-in for the LDGSTS gather writes)
__global__ void __launch_bounds__(256, 1) k_probe(int n, int iters, int mode, int writers, long long *out, const uint8_t *src, int wmode) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t done;
    __shared__ __align__(8) uint64_t ring[8];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    // A: 128 rows x 128 B (64 bf16) = 16 KB; B: 256 rows x 128 B = 32 KB; scratch behind
    for (int i = tid; i < (16 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        mbar_init(smem_u32(&done), 1);
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&ring[i]), 1);
        mbar_fence_init();
        stop = 0;
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 16 * 1024, scratch = b_base + 32 * 1024;
    if (warp == 0) {
        const uint32_t idesc = idesc_bf16(n);
        const uint64_t da = desc_k_sw128(a_base), db = desc_k_sw128(b_base);
        long long t0 = clock64();
        if (wmode == 4) {
            // mode = MMAs per commit (0: commits only, no MMA at all)
            const int g = mode;
            if (elect_one()) {
                for (int i = 0; i < iters; ++i) {
                    if (g > 0) umma(tmem_base, da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), idesc, i >= 2);
                    if (g == 0 || (i % g) == g - 1) umma_commit(smem_u32(&ring[(i / (g ? g : 1)) & 7]));
                }
                umma_commit(smem_u32(&done));
            }
        } else if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
                const uint64_t adv = mode == 2 ? (uint64_t)((i & 3) * 2) : 0;
                const uint32_t acc = tmem_base + ((mode == 1 && (i & 1)) ? 256u : 0u);
                umma(acc, da + adv, db + adv, idesc, i >= 2);
            }
            umma_commit(smem_u32(&done));
        }
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(smem_u32(&done), 0, 0);
        long long t2 = clock64();
        stop = 1;
        if (tid == 0) {
            out[blockIdx.x * 2] = t1 - t0;
            out[blockIdx.x * 2 + 1] = t2 - t0;
        }
    } else if (wmode < 2 && warp <= writers) {
        uint32_t dst = scratch + (uint32_t)(warp - 1) * 4096u + (uint32_t)(tid & 31) * 16u;
        int k = 0;
        long long bytes = 0;
        if (wmode == 0) {
            while (!stop) {
                asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst + (uint32_t)((k & 7) * 512)), "r"(k) : "memory");
                ++k;
            }
        } else {
            // the conv kernel's gather: 8 threads per 128-byte row, XOR-swizzled 16-byte chunks, rows scattered over a 32 MB buffer
            const int lane = tid & 31, grp = lane & 7, r4 = lane >> 3;
            unsigned h = (unsigned)(blockIdx.x * 977 + warp * 131 + 7);
            while (!stop) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    h = h * 1664525u + 1013904223u;
                    const unsigned row = (h >> 8) & 0x3FFFFu;  // 262144 rows x 128 B
                    const int r = i * 4 + r4;
                    cp_async16(scratch + (uint32_t)(warp - 1) * 4096u + (uint32_t)r * 128u + (uint32_t)((grp ^ (r & 7)) << 4),
                               src + (size_t)row * 128 + grp * 16, 16u);
                }
                cp_async_commit();
                cp_async_wait<2>();
                bytes += 8 * 512;
                ++k;
            }
            cp_async_wait<0>();
            if (lane == 0) atomicAdd((unsigned long long *)&out[296 + blockIdx.x], (unsigned long long)bytes);
        }
    }
    else if (wmode >= 2 && warp >= 1) {
        // wmode 2: all 32 lanes of 7 warps poll an mbarrier that never completes (what the conv kernel's waiting roles do);
        // wmode 3: one lane polls, the others park at __syncwarp
        __shared__ __align__(8) uint64_t never;
        if (tid == 32) mbar_init(smem_u32(&never), 1);
        __syncwarp();
        if (wmode == 2 || (tid & 31) == 0)
            while (!stop) mbar_try_wait(smem_u32(&never), 0);
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
    long long *out;
    cudaMalloc(&out, 148 * 3 * sizeof(long long));
    uint8_t *src;
    cudaMalloc(&src, 32 << 20);
    cudaMemset(src, 1, 32 << 20);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4096;
    long long h[444];
    for (int wmode = 0; wmode < 5; ++wmode)
    for (int writers = (wmode == 1 ? 2 : 0); writers <= (wmode == 1 ? 7 : wmode == 0 ? 4 : 0); writers += (wmode == 1 ? 1 : 4))
        for (int mode = ((wmode && wmode < 4) ? 2 : 0); mode < (wmode == 4 ? 9 : 3); ++mode)
            for (int n : {32, 64, 96, 128, 192, 256}) {
                if (wmode && n != 96 && n != 192 && n != 256) continue;
                if (wmode == 4 && n != 96) continue;
                if (wmode == 4 && !(mode == 0 || mode == 1 || mode == 2 || mode == 4 || mode == 8)) continue;
                cudaMemset(out, 0, 148 * 3 * sizeof(long long));
                k_probe<<<148, 256, 200 * 1024>>>(n, iters, mode, writers, out, src, wmode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                double issue = 0, total = 0, wb = 0;
                for (int b = 0; b < 148; ++b) { issue += h[2 * b]; total += h[2 * b + 1]; wb += h[296 + b]; }
                printf("%s writers %d mode %d N %3d: %.1f cyc/MMA issue, %.1f cyc/MMA complete (floor 128*N/256 = %d), gather %.1f B/clk/SM\n",
                       wmode == 0 ? "st.shared" : wmode == 1 ? "cp.async" : wmode == 2 ? "7 warps x 32 lanes polling" : wmode == 3 ? "7 warps x 1 lane polling" : "commit every <mode> MMAs", writers, mode, n, issue / 148 / iters, total / 148 / iters, n / 2, wb / total);
            }
    return 0;
}
