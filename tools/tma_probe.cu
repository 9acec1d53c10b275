// Standalone probe: which fp32 TMA box shapes are accepted?  usage: tma_probe W H C bx by bz cx cy cz
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int n, int cx, int cy, int cz) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar), dst = (unsigned)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(n * 4));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"((unsigned long long)&tm), "r"(bar_a), "r"(cx), "r"(cy), "r"(cz) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(bar_a));
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = ((float*)sm)[i];
}
int main(int argc, char** argv) {
    int W = atoi(argv[1]), H = atoi(argv[2]), C = atoi(argv[3]), bx = atoi(argv[4]), by = atoi(argv[5]), bz = atoi(argv[6]);
    int cx = atoi(argv[7]), cy = atoi(argv[8]), cz = atoi(argv[9]);
    void* sym; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    enc_fn enc = (enc_fn)sym;
    std::vector<float> h(W * H * C);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int n = bx * by * bz; cudaMalloc(&o, n * 4);
    CUtensorMap tm; cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C}, str[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%dx%d at (%d,%d,%d): encode=%d ", bx, by, bz, cx, cy, cz, (int)r);
    if (r) { printf("\n"); return 0; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<<<1, 128, n * 4 + 1024>>>(tm, o, n, cx, cy, cz);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s", cudaGetErrorString(e));
    if (!e) {
        std::vector<float> ho(n); cudaMemcpy(ho.data(), o, n * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int z = 0; z < bz; ++z) for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
            int gx = cx + x, gy = cy + y, gz = cz + z;
            float want = (gx < 0 || gx >= W || gy < 0 || gy >= H || gz < 0 || gz >= C) ? 0.f : (float)((gz * H + gy) * W + gx);
            if (ho[(z * by + y) * bx + x] != want) ++bad;
        }
        printf(" mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
