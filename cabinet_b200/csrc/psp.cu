// Pyramid pooling pieces (src/models/cab.py:46-76): adaptive average pools (1,3,6,8) and the
// [x, up(pool_s)...] 5C-channel concat that feeds the 1x1 `project` GEMM.  The maps here are 1/32 resolution
// (<= 8160 pixels, 128 channels): negligible bytes next to the high-resolution layers.
#include "common.cuh"

namespace {

__device__ __constant__ int kPspSize[4] = {1, 3, 6, 8};
__device__ __constant__ int kPspOffset[4] = {0, 1, 10, 46};  // bin offsets inside the 110-entry table
constexpr int kPspBins = 110;

// grid (110, N): one bin per block; threads = (C/4 channel quads) x (pixel lanes), smem tree over the pixel lanes.
template <typename T>
__global__ void __launch_bounds__(1024)
psp_pool_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ pooled, int H, int W, int C) {
    __shared__ float4 red[1024];
    const int bin = blockIdx.x, n = blockIdx.y;
    int si = 3;
    if (bin < 1) si = 0; else if (bin < 10) si = 1; else if (bin < 46) si = 2;
    const int s = kPspSize[si];
    const int local = bin - kPspOffset[si];
    const int by = local / s, bx = local % s;
    // ATen adaptive_avg_pool2d bins: [floor(i*H/s), ceil((i+1)*H/s))
    const int h0 = (by * H) / s, h1 = ((by + 1) * H + s - 1) / s;
    const int w0 = (bx * W) / s, w1 = ((bx + 1) * W + s - 1) / s;
    const int bw = w1 - w0, npix = (h1 - h0) * bw;
    const int CQ = C / 4;                    // channel quads (C % 4 == 0, CQ <= 256 checked by the host)
    const int lanes = blockDim.x / CQ;       // pixel lanes
    const int cq = threadIdx.x % CQ, pl = threadIdx.x / CQ;
    const T* base = x + static_cast<long long>(n) * H * W * ldx + cq * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pl < lanes) {
        for (int i = pl; i < npix; i += lanes) {
            const int h = h0 + i / bw, w = w0 + i % bw;
            const T* p = base + (static_cast<long long>(h) * W + w) * ldx;
            acc.x += to_f32<T>(p[0]); acc.y += to_f32<T>(p[1]); acc.z += to_f32<T>(p[2]); acc.w += to_f32<T>(p[3]);
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (pl == 0) {
        for (int l = 1; l < lanes; ++l) {
            const float4 o = red[l * CQ + cq];
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        const float inv = 1.f / static_cast<float>(npix);
        float* out = pooled + (static_cast<long long>(n) * kPspBins + bin) * C + cq * 4;
        *reinterpret_cast<float4*>(out) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
}

// grid (ceil(W*C/256), H, N); thread per (column, channel): out[pix][0:C] = x, out[pix][C*(1+si) + c] =
// bilinear(pooled_si)(pix, c).  32-bit index arithmetic; row taps are per-block constants.
template <typename T>
__global__ void __launch_bounds__(256)
psp_concat_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ pooled, T* __restrict__ out,
                  long long ldo, int H, int W, int C) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= static_cast<unsigned>(W) * C) return;
    const int w = idx / C, c = idx - w * C;
    const int h = blockIdx.y, n = blockIdx.z;
    const long long pix = (static_cast<long long>(n) * H + h) * W + w;
    T* o = out + pix * ldo;
    o[c] = x[pix * ldx + c];
    const float* pn = pooled + static_cast<long long>(n) * kPspBins * C;
#pragma unroll
    for (int si = 0; si < 4; ++si) {
        const int s = kPspSize[si];
        const float* ps = pn + static_cast<long long>(kPspOffset[si]) * C;
        int y0, y1, x0, x1;
        float wy, wx;
        cab_bilinear_tap(h, static_cast<float>(s) / static_cast<float>(H), s, y0, y1, wy);
        cab_bilinear_tap(w, static_cast<float>(s) / static_cast<float>(W), s, x0, x1, wx);
        const float v00 = ps[(y0 * s + x0) * C + c], v01 = ps[(y0 * s + x1) * C + c];
        const float v10 = ps[(y1 * s + x0) * C + c], v11 = ps[(y1 * s + x1) * C + c];
        const float v = (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11);
        o[static_cast<long long>(C) * (1 + si) + c] = from_f32<T>(v);
    }
}

}  // namespace

extern "C" int cabinet_psp_pool(const void* x, long long ldx, int dtype, float* pooled, int N, int H, int W, int C,
                                cabinet_stream_t stream) {
    CAB_REQUIRE(x && pooled && H > 0 && W > 0 && C > 0 && ldx >= C && C % 4 == 0 && C <= 1024,
                "psp_pool: bad arguments (C must be a multiple of 4, <= 1024)");
    if (N == 0) return CABINET_OK;
    dim3 grid(kPspBins, N);
    const int threads = 1024;  // (C/4 channel quads) x (pixel lanes): the s = 1 bin spans the whole map
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == CABINET_BF16)
        psp_pool_kernel<bf16><<<grid, threads, 0, s>>>(reinterpret_cast<const bf16*>(x), ldx, pooled, H, W, C);
    else
        psp_pool_kernel<float><<<grid, threads, 0, s>>>(reinterpret_cast<const float*>(x), ldx, pooled, H, W, C);
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}

extern "C" int cabinet_psp_concat(const void* x, long long ldx, const float* pooled, void* out, long long ldo,
                                  int dtype, int N, int H, int W, int C, cabinet_stream_t stream) {
    CAB_REQUIRE(x && pooled && out && H > 0 && W > 0 && C > 0 && ldx >= C && ldo >= 5LL * C,
                "psp_concat: bad arguments");
    if (N == 0) return CABINET_OK;
    CAB_REQUIRE(H <= 65535 && N <= 65535, "psp_concat: H/N exceed grid limits");
    dim3 grid(static_cast<unsigned>(cab_ceil_div(static_cast<long long>(W) * C, 256)), H, N);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == CABINET_BF16)
        psp_concat_kernel<bf16><<<grid, 256, 0, s>>>(reinterpret_cast<const bf16*>(x), ldx, pooled,
                                                    reinterpret_cast<bf16*>(out), ldo, H, W, C);
    else
        psp_concat_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(x), ldx, pooled,
                                                     reinterpret_cast<float*>(out), ldo, H, W, C);
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}
