// Per-kernel latency of a chain of dependent small kernels inside a captured CUDA graph, with / without PDL.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k(float* x, int n, int pdl) {
    __shared__ float s[1024];
    s[threadIdx.x] = threadIdx.x;          // "prologue": independent of the previous kernel
    __syncthreads();
    if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = x[i] * 1.0001f + s[(threadIdx.x + 1) & 1023] * 1e-9f;
}
static void launch(float* x, int n, int grid, int pdl, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(1024); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k, x, n, pdl);
}
int main() {
    const int n = 148 * 1024, chain = 200;
    float* x; cudaMalloc(&x, n * 4); cudaMemset(x, 0, n * 4);
    cudaStream_t st; cudaStreamCreate(&st);
    for (int grid : {148, 148 * 8}) for (int pdl = 0; pdl < 2; ++pdl) {
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        for (int i = 0; i < chain; ++i) launch(x, n, grid, pdl, st);
        cudaError_t e1 = cudaStreamEndCapture(st, &g);
        cudaError_t e2 = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st); for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, st); cudaEventRecord(b, st);
        cudaError_t e3 = cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("grid %5d pdl %d: %.2f us per kernel (capture %s, instantiate %s, run %s)\n", grid, pdl, ms * 1e3 / (5 * chain),
               cudaGetErrorString(e1), cudaGetErrorString(e2), cudaGetErrorString(e3));
    }
    return 0;
}
