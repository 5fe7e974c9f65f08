// Hand-off latency inside one CTA: two warps play ping-pong N times through
//   mode 0: two mbarriers (mbarrier.arrive by lane 0 / all 32 lanes poll with mbarrier.try_wait.parity)
//   mode 1: the same with 128 arriving threads on the ping side (4 warps, like the conv kernel's producers: count 128)
//   mode 2: volatile shared-memory flags (st.volatile + ld.volatile spin)
//   mode 3: ping by tcgen05.commit after one M128 N96 K16 MMA (the consumer -> producer direction of the conv kernel)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I unscene3d_b200/csrc scripts/experiments/mbar_pingpong_probe.cu -o scripts/mbar_pingpong_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace us3d::tcx;

__global__ void __launch_bounds__(192, 1) k_probe(int mode, int rounds, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t ping, pong;
    __shared__ volatile int f_ping, f_pong;
    __shared__ uint32_t tmem_base_s;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 32 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        mbar_init(smem_u32(&ping), mode == 1 ? 128 : 1);
        mbar_init(smem_u32(&pong), 1);
        mbar_fence_init();
        f_ping = f_pong = 0;
    }
    if (warp == 4) tmem_alloc(&tmem_base_s, 128);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 16 * 1024;
    long long t0 = clock64();
    if (warp == 4) {  // side A: sends ping, waits pong
        for (int r = 0; r < rounds; ++r) {
            if (mode == 0) {
                if (lane == 0) mbar_arrive(smem_u32(&ping));
                __syncwarp();
                mbar_wait(smem_u32(&pong), r & 1, 0);
            } else if (mode == 2) {
                if (lane == 0) f_ping = r + 1;
                while (f_pong != r + 1) {
                }
                __syncwarp();
            } else if (mode == 3) {
                if (elect_one()) {
                    umma(tmem_base_s, desc_k_sw128(a_base), desc_k_sw128(b_base), idesc_bf16(96), 0);
                    umma_commit(smem_u32(&ping));
                }
                __syncwarp();
                mbar_wait(smem_u32(&pong), r & 1, 0);
            } else {
                mbar_wait(smem_u32(&pong), r & 1, 0);
            }
        }
        if (lane == 0) out[blockIdx.x] = clock64() - t0;
    } else if (warp == 5) {  // side B: waits ping, sends pong
        for (int r = 0; r < rounds; ++r) {
            if (mode == 2) {
                while (f_ping != r + 1) {
                }
                __syncwarp();
                if (lane == 0) f_pong = r + 1;
            } else {
                mbar_wait(smem_u32(&ping), r & 1, 1);
                if (lane == 0) mbar_arrive(smem_u32(&pong));
                __syncwarp();
            }
        }
    } else if (mode == 1) {  // warps 0-3: 128 arriving threads, paced by pong like the producers by a freed slot
        for (int r = 0; r < rounds; ++r) {
            mbar_arrive(smem_u32(&ping));
            mbar_wait(smem_u32(&pong), r & 1, 2);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base_s, 128);
}

int main() {
    long long *out;
    cudaMalloc(&out, 148 * sizeof(long long));
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int rounds = 2000;
    const char *names[] = {"mbarrier, 1 arrival", "mbarrier, 128 arriving threads", "volatile shared flags", "tcgen05.commit after one MMA"};
    for (int mode = 0; mode < 4; ++mode) {
        k_probe<<<148, 192, 64 * 1024>>>(mode, rounds, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int b = 0; b < 148; ++b) s += h[b];
        printf("%-34s %.0f cycles per round trip (two hand-offs)\n", names[mode], s / 148 / rounds);
    }
    return 0;
}
