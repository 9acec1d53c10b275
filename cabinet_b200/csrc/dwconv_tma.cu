// Depthwise k x k convolution, bf16 NHWC, shared-memory halo staging by TMA.
//
// One CTA = one 8 x 16 output patch of one image for one chunk of <= 64 channels.  A single 4-D TMA box
// {CB channels, (16-1)*S+K cols, (8-1)*S+K rows, 1} brings the input patch INCLUDING its halo into shared memory;
// the zero padding of the convolution is the TMA out-of-bounds fill, so the compute loop has no bounds checks and no
// global loads except the (L1-resident) weights.  Each thread owns 8 channels x 4 consecutive output columns and slides
// a register window along the row (K+3*S 16-byte LDS per filter row).  Several CTAs are resident per SM, so the TMA
// fetch of one overlaps the FMAs / stores of the others.  Outputs are written with 16-byte stores that cover whole
// 128-byte lines (8 lanes = the 64 channels of one pixel).
// Epilogue: + folded-BN bias, activation, optional per-(image, channel) partial sums for the SE / GAP consumers.
// HBM-bound: algorithmic bytes = input + output (+ weights); halo re-reads are served by L2.
#include "tc_common.cuh"

namespace {

constexpr int TH = 8, TW = 16, RUN = 4;

struct DwParams {
    int C, OH, OW, CB, act;
    const float* w;
    const float* bias;
    bf16* y;
    long long ldy;
    float* gap_sum;
};

template <int K, int S>
__global__ void __launch_bounds__(256)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmX, const DwParams p) {
    constexpr int PAD = (K - 1) / 2;
    constexpr int IWT = (TW - 1) * S + K, IHT = (TH - 1) * S + K;
    constexpr int SPAN = (RUN - 1) * S + K;
    extern __shared__ __align__(128) uint8_t smem_dw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_gap[64];

    const int tiles_w = (p.OW + TW - 1) / TW;
    const int ow0 = (blockIdx.x % tiles_w) * TW, oh0 = (blockIdx.x / tiles_w) * TH;
    const int chunk = blockIdx.y, n = blockIdx.z;
    const int CB = p.CB, CGB = CB >> 3;
    uint8_t* tile = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dw) + 127) & ~uintptr_t(127));

    if (threadIdx.x == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_fence_init();
        tc::fence_proxy_async();
        tc::mbar_expect_tx(&bar, static_cast<uint32_t>(IWT * IHT * CB * 2));
        tc::tma_load_4d(tile, &tmX, &bar, chunk * CB, ow0 * S - PAD, oh0 * S - PAD, n);
    }
    if (p.gap_sum && threadIdx.x < 64) s_gap[threadIdx.x] = 0.f;
    __syncthreads();

    const int cg = threadIdx.x % CGB;
    const int rem = threadIdx.x / CGB;
    const int run = rem % (TW / RUN), row = rem / (TW / RUN);
    const int c0 = chunk * CB + cg * 8;
    const bool active = row < TH && c0 < p.C;  // C % 8 == 0, so a group is either fully inside or outside

    float acc[RUN][8];
    if (active) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + 1);
#pragma unroll
        for (int r = 0; r < RUN; ++r) {
            acc[r][0] = b0.x; acc[r][1] = b0.y; acc[r][2] = b0.z; acc[r][3] = b0.w;
            acc[r][4] = b1.x; acc[r][5] = b1.y; acc[r][6] = b1.z; acc[r][7] = b1.w;
        }
    }
    tc::mbar_wait(&bar, 0);
    if (active) {
        const int pitch = CB * 2;  // bytes per staged pixel
        const uint8_t* base = tile + ((row * S) * IWT + run * RUN * S) * pitch + cg * 16;
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            float wk[K][8];
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w + (ky * K + kx) * p.C + c0));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w + (ky * K + kx) * p.C + c0) + 1);
                wk[kx][0] = w0.x; wk[kx][1] = w0.y; wk[kx][2] = w0.z; wk[kx][3] = w0.w;
                wk[kx][4] = w1.x; wk[kx][5] = w1.y; wk[kx][6] = w1.z; wk[kx][7] = w1.w;
            }
            const uint8_t* rowp = base + ky * IWT * pitch;
#pragma unroll
            for (int sx = 0; sx < SPAN; ++sx) {
                Vec16<bf16> xv;
                xv.raw = *reinterpret_cast<const uint4*>(rowp + sx * pitch);
                float xf[8];
                xv.unpack(xf);
#pragma unroll
                for (int r = 0; r < RUN; ++r) {
                    const int kx = sx - r * S;  // compile-time after unrolling
                    if (kx >= 0 && kx < K) {
#pragma unroll
                        for (int v = 0; v < 8; ++v) acc[r][v] = fmaf(xf[v], wk[kx][v], acc[r][v]);
                    }
                }
            }
        }
        const int oh = oh0 + row;
        float gsum[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) gsum[v] = 0.f;
        if (oh < p.OH) {
            bf16* yout = p.y + ((static_cast<long long>(n) * p.OH + oh) * p.OW) * p.ldy + c0;
#pragma unroll
            for (int r = 0; r < RUN; ++r) {
                const int ow = ow0 + run * RUN + r;
                if (ow >= p.OW) continue;
                cab_act_vec<8>(acc[r], p.act);
#pragma unroll
                for (int v = 0; v < 8; ++v) gsum[v] += acc[r][v];
                Vec16<bf16> ov;
                ov.pack(acc[r]);
                ov.store(yout + static_cast<long long>(ow) * p.ldy);
            }
        }
        if (p.gap_sum) {
#pragma unroll
            for (int v = 0; v < 8; ++v) atomicAdd(&s_gap[cg * 8 + v], gsum[v]);
        }
    }
    if (p.gap_sum) {
        __syncthreads();
        const int c = chunk * CB + threadIdx.x;
        if (threadIdx.x < CB && c < p.C) atomicAdd(&p.gap_sum[static_cast<long long>(n) * p.C + c], s_gap[threadIdx.x]);
    }
}

template <int K, int S>
int launch(const CUtensorMap& tm, const DwParams& p, int N, int n_chunks, cudaStream_t s) {
    constexpr int IWT = (TW - 1) * S + K, IHT = (TH - 1) * S + K;
    const size_t smem = static_cast<size_t>(IWT) * IHT * p.CB * 2 + 128;
    static bool attr_set = false;
    if (!attr_set) {
        CAB_CUDA(cudaFuncSetAttribute(dwconv_tma_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    const int tiles = ((p.OW + TW - 1) / TW) * ((p.OH + TH - 1) / TH);
    dim3 grid(tiles, n_chunks, N);
    const int threads = (p.CB / 8) * (TW / RUN) * TH;  // 32 * CB/8 <= 256
    dwconv_tma_kernel<K, S><<<grid, threads, smem, s>>>(tm, p);
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}

}  // namespace

extern "C" int cabinet_dwconv_tma(const void* x, long long ldx, const float* w, const float* bias, void* y,
                                  long long ldy, int N, int H, int W, int C, int k, int stride, int OH, int OW, int act,
                                  float* gap_sum, cabinet_stream_t stream) {
    CAB_REQUIRE(x && w && bias && y, "dwconv_tma: null pointer");
    CAB_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2), "dwconv_tma: k must be 3|5 and stride 1|2");
    CAB_REQUIRE(C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C &&
                    (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                "dwconv_tma: C/ldx/ldy must be multiples of 8 and pointers 16-byte aligned");
    CAB_REQUIRE(OH == (H + 2 * ((k - 1) / 2) - k) / stride + 1 && OW == (W + 2 * ((k - 1) / 2) - k) / stride + 1,
                "dwconv_tma: inconsistent output size");
    CAB_REQUIRE(N <= 65535, "dwconv_tma: N exceeds grid limits");
    if (N == 0) return CABINET_OK;
    DwParams p;
    const int n_chunks = (C + 63) / 64;
    p.C = C; p.OH = OH; p.OW = OW; p.act = act; p.w = w; p.bias = bias;
    p.CB = ((C + n_chunks - 1) / n_chunks + 7) / 8 * 8;
    p.y = reinterpret_cast<bf16*>(y); p.ldy = ldy; p.gap_sum = gap_sum;
    const int IWT = (TW - 1) * stride + k, IHT = (TH - 1) * stride + k;
    CUtensorMap tm;
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * W, (uint64_t)ldx * 2 * W * H};
    const uint32_t box[4] = {(uint32_t)p.CB, (uint32_t)IWT, (uint32_t)IHT, 1};
    int rc = cab_make_tmap_bf16(&tm, x, 4, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (k == 3 && stride == 1) return launch<3, 1>(tm, p, N, n_chunks, s);
    if (k == 3 && stride == 2) return launch<3, 2>(tm, p, N, n_chunks, s);
    if (k == 5 && stride == 1) return launch<5, 1>(tm, p, N, n_chunks, s);
    return launch<5, 2>(tm, p, N, n_chunks, s);
}
