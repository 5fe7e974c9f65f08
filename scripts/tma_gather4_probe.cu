// Stand-alone probe: how does cp.async.bulk.tensor.2d...tile::gather4 place 4 gathered rows in shared memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_probe scripts/tma_gather4_probe.cu
// Matrix bf16 [R][C], value(r, c) = r * 4 + c / 32 (exact in bf16 for small r); gathers rows {5, 17, -1, R-1, ...}.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int col, int r0, int r1, int r2, int r3, uint16_t *out, int nbytes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    for (int i = threadIdx.x; i < nbytes / 2; i += blockDim.x) reinterpret_cast<uint16_t *>(smem)[i] = 0x7FC0;  // NaN marker
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(4 * 128) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(dst), "l"(&tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_a)
            : "memory");
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok)
                         : "r"(bar_a)
                         : "memory");
            if (clock64() - t0 > 2000000000LL) {
                printf("timeout waiting for gather4\n");
                break;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t *>(smem)[i];
}

static float bf(uint16_t v) {
    uint32_t u = (uint32_t)v << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int main() {
    const int R = 1000, C = 128;
    std::vector<__nv_bfloat16> h((size_t)R * C);
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = __float2bfloat16((float)(r % 60) * 4 + (float)(c / 8) / 8.f);
    __nv_bfloat16 *d;
    cudaMalloc(&d, h.size() * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &q);
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    uint16_t *out;
    const int nbytes = 2048;
    cudaMalloc(&out, nbytes);
    for (int box_rows = 1; box_rows <= 4; box_rows += 3) {
        CUtensorMap tmap;
        cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)R};
        cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult rc = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box {64,%d}: encode rc=%d\n", box_rows, (int)rc);
        if (rc != CUDA_SUCCESS) continue;
        for (int col = 0; col <= 64; col += 64) {
            cudaMemset(out, 0, nbytes);
            probe<<<1, 128, nbytes + 1024>>>(tmap, col, 5, 17, -1, R - 1, out, nbytes);
            cudaError_t e = cudaDeviceSynchronize();
            printf(" col %d launch: %s\n", col, cudaGetErrorString(e));
            if (e != cudaSuccess) return 2;
            std::vector<uint16_t> o(nbytes / 2);
            cudaMemcpy(o.data(), out, nbytes, cudaMemcpyDeviceToHost);
            // print first element of every 16-byte chunk for the first 8 rows of 128 B
            for (int row = 0; row < 8; ++row) {
                printf("  smem row %d:", row);
                for (int ch = 0; ch < 8; ++ch) printf(" %7.3f", bf(o[row * 64 + ch * 8]));
                printf("\n");
            }
            printf("  expected values: row r -> (r%%60)*4 + chunk/8 (+1.0 for col 64): rows 5->20.x, 17->68.x, -1->0, 999->156.x\n");
        }
    }
    return 0;
}
