// Cost of tcgen05.commit for the MMA-issuing thread: G MMAs (M128 x N x K16, SS mode) then one commit onto a rotating mbarrier,
// repeated; compile-time G and N.  Also: the same with a consumer warp that waits for every commit (like a producer waiting for
// a freed ring slot).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I unscene3d_b200/csrc scripts/experiments/umma_commit_probe.cu -o scripts/umma_commit_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace us3d::tcx;

template <int G, int N, bool WAITER>
__global__ void __launch_bounds__(128, 1) k_probe(int groups, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t done, ring[8];
    __shared__ uint32_t tmem_base_s;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        mbar_init(smem_u32(&done), 1);
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&ring[i]), 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 16 * 1024;
    if (warp == 0) {
        const uint32_t idesc = idesc_bf16(N);
        const uint64_t da = desc_k_sw128(a_base), db = desc_k_sw128(b_base);
        long long t0 = clock64();
        int slot = 0;
        for (int g = 0; g < groups; ++g) {
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < G; ++i) umma(tmem_base, da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), idesc, 1);
                umma_commit(smem_u32(&ring[slot]));
            }
            __syncwarp();
            slot = (slot + 1) & 7;
        }
        if (elect_one()) umma_commit(smem_u32(&done));
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(smem_u32(&done), 0, 0);
        long long t2 = clock64();
        if (tid == 0) {
            out[blockIdx.x * 2] = t1 - t0;
            out[blockIdx.x * 2 + 1] = t2 - t0;
        }
    } else if (WAITER && warp == 1) {
        int slot = 0;
        uint32_t par = 0;
        for (int g = 0; g < groups; ++g) {
            mbar_wait(smem_u32(&ring[slot]), par, 1);
            if (++slot == 8) { slot = 0; par ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int G, int N, bool WAITER>
void run(long long *out) {
    cudaFuncSetAttribute(k_probe<G, N, WAITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int groups = 4096 / (G ? G : 1);
    k_probe<G, N, WAITER><<<148, 128, 64 * 1024>>>(groups, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[296];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double issue = 0, total = 0;
    for (int b = 0; b < 148; ++b) { issue += h[2 * b]; total += h[2 * b + 1]; }
    printf("G %2d N %3d waiter %d: %.1f cyc/group issue, %.1f cyc/group complete; MMA alone would be %d\n", G, N, (int)WAITER,
           issue / 148 / groups, total / 148 / groups, G * (N <= 128 ? 32 + N / 4 : N / 2));
}

int main() {
    long long *out;
    cudaMalloc(&out, 296 * sizeof(long long));
    run<0, 96, false>(out); run<1, 96, false>(out); run<2, 96, false>(out); run<4, 96, false>(out); run<8, 96, false>(out); run<16, 96, false>(out);
    run<1, 96, true>(out); run<4, 96, true>(out); run<8, 96, true>(out);
    run<4, 192, false>(out); run<8, 192, false>(out); run<4, 256, false>(out);
    return 0;
}
