// Pure TMA-load bandwidth probe: persistent CTAs stream [rows x 64 bf16] boxes of a [P][C] bf16 tensor through an
// smem ring; a consumer thread only recycles the slots.  usage: tma_bw C box_rows stages ctas_per_sm [swizzle128=1]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(s32(b)), "r"(ph) : "memory");
}
__global__ void __launch_bounds__(64) k(const __grid_constant__ CUtensorMap tm, int box_rows, int stages, int tiles, int kblocks) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t full[16], empty[16];
    uint8_t* base = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = box_rows * 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int g = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x)
            for (int kb = 0; kb < kblocks; ++kb, ++g) {
                int s = g % stages; uint32_t ph = (g / stages) & 1;
                mwait(&empty[s], ph ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(stage_bytes) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(s32(base + s * stage_bytes)), "l"((uint64_t)&tm), "r"(s32(&full[s])), "r"(kb * 64), "r"(t * box_rows) : "memory");
            }
    } else if (threadIdx.x == 32) {
        int g = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x)
            for (int kb = 0; kb < kblocks; ++kb, ++g) {
                int s = g % stages; uint32_t ph = (g / stages) & 1;
                mwait(&full[s], ph);
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
            }
    }
}
int main(int argc, char** argv) {
    int C = atoi(argv[1]), box_rows = atoi(argv[2]), stages = atoi(argv[3]), cps = atoi(argv[4]);
    int sw = argc > 5 ? atoi(argv[5]) : 1;
    void* sym; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    enc_fn enc = (enc_fn)sym;
    const long long P = (1LL << 30) / (C * 2);  // 1 GiB tensor
    void* d; cudaMalloc(&d, P * C * 2); cudaMemset(d, 0, P * C * 2);
    CUtensorMap tm; cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)P}, str[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode failed %d\n", (int)r); return 1; }
    const int tiles = (int)(P / box_rows), kblocks = (C + 63) / 64;
    size_t smem = (size_t)stages * box_rows * 128 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * cps, 64, smem>>>(tm, box_rows, stages, tiles, kblocks);
    cudaEventRecord(e0);
    k<<<148 * cps, 64, smem>>>(tm, box_rows, stages, tiles, kblocks);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)P * (C < 64 ? C : C) * 2;
    printf("C=%4d box_rows=%3d stages=%d ctas/sm=%d sw=%d: %s %.1f us  %.0f GB/s (%.2f us per box per SM)\n", C, box_rows, stages, cps, sw,
           cudaGetErrorString(e), ms * 1e3, bytes / ms / 1e6, ms * 1e3 / ((double)tiles * kblocks / 148));
    return 0;
}
