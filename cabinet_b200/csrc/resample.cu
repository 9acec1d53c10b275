// Bilinear resampling (align_corners=False) and the fused output tail:
//   bilinear_nhwc          1/32 -> 1/8 feature / aux-logit upsample (src/models/cabinet.py:228-233)
//   upsample_logits_nchw   x8 logits upsample written as the NCHW tensors CABiNet.forward returns (:240-245)
//   upsample_argmax        the same upsample fused with argmax (+ confusion matrix), logits never reach HBM
//   confusion_hist         hist[pred, label] (src/scripts/evaluate.py:162-191)
// All HBM-bound on their OUTPUT bytes: the sources are <= 1/64 of the output and stay in L1/L2.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------- NHWC -> NHWC
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
bilinear_nhwc_kernel(const TI* __restrict__ x, long long ldx, TO* __restrict__ y, long long ldy, int IH, int IW, int C,
                     int OH, int OW, long long total, float sh, float sw) {
    // thread per (output pixel, group of 4 channels)
    const int CG = (C + 3) / 4;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cg = static_cast<int>(idx % CG);
    long long t = idx / CG;
    const int ow = static_cast<int>(t % OW);
    t /= OW;
    const int oh = static_cast<int>(t % OH);
    const int n = static_cast<int>(t / OH);
    int y0, y1, x0, x1;
    float wy, wx;
    cab_bilinear_tap(oh, sh, IH, y0, y1, wy);
    cab_bilinear_tap(ow, sw, IW, x0, x1, wx);
    const TI* b = x + static_cast<long long>(n) * IH * IW * ldx + cg * 4;
    const TI* p00 = b + (static_cast<long long>(y0) * IW + x0) * ldx;
    const TI* p01 = b + (static_cast<long long>(y0) * IW + x1) * ldx;
    const TI* p10 = b + (static_cast<long long>(y1) * IW + x0) * ldx;
    const TI* p11 = b + (static_cast<long long>(y1) * IW + x1) * ldx;
    TO* o = y + ((static_cast<long long>(n) * OH + oh) * OW + ow) * ldy + cg * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (cg * 4 + c >= C) break;
        const float v = (1.f - wy) * ((1.f - wx) * to_f32<TI>(p00[c]) + wx * to_f32<TI>(p01[c])) +
                        wy * ((1.f - wx) * to_f32<TI>(p10[c]) + wx * to_f32<TI>(p11[c]));
        o[c] = from_f32<TO>(v);
    }
}

// ---------------------------------------------------------------------------------- logits tail
// Each thread produces PX consecutive output pixels of one row for ALL classes.
// Source rows are addressed through the two y taps; x taps are recomputed per pixel (cheap).


template <typename TO, int PX>
__global__ void __launch_bounds__(256)
upsample_logits_nchw_kernel(const float* __restrict__ x, int IH, int IW, int C, TO* __restrict__ y, int OH, int OW,
                            float sh, float sw) {
    const int groups_w = (OW + PX - 1) / PX;
    const int gw = blockIdx.x * blockDim.x + threadIdx.x;
    if (gw >= groups_w) return;
    const int oh = blockIdx.y, n = blockIdx.z;
    int y0, y1;
    float wy;
    cab_bilinear_tap(oh, sh, IH, y0, y1, wy);
    const float* r0 = x + (static_cast<long long>(n) * IH + y0) * IW * C;
    const float* r1 = x + (static_cast<long long>(n) * IH + y1) * IW * C;
    const int ow0 = gw * PX;
    int x0[PX], x1[PX];
    float wx[PX];
#pragma unroll
    for (int p = 0; p < PX; ++p) cab_bilinear_tap(min(ow0 + p, OW - 1), sw, IW, x0[p], x1[p], wx[p]);
    const bool full = (ow0 + PX <= OW) && (OW % PX == 0);
    for (int c = 0; c < C; ++c) {
        float v[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const float a = (1.f - wx[p]) * __ldg(r0 + x0[p] * C + c) + wx[p] * __ldg(r0 + x1[p] * C + c);
            const float b = (1.f - wx[p]) * __ldg(r1 + x0[p] * C + c) + wx[p] * __ldg(r1 + x1[p] * C + c);
            v[p] = (1.f - wy) * a + wy * b;
        }
        TO* o = y + ((static_cast<long long>(n) * C + c) * OH + oh) * OW + ow0;
        if (full) {
            if constexpr (sizeof(TO) == 4) {
#pragma unroll
                for (int p = 0; p < PX; p += 4)
                    __stcs(reinterpret_cast<float4*>(o + p), make_float4(v[p], v[p + 1], v[p + 2], v[p + 3]));
            } else {
                static_assert(PX == 8, "bf16 path stores 8 pixels = 16 bytes");
                Vec16<bf16> ov;
                ov.pack(v);
                __stcs(reinterpret_cast<uint4*>(o), ov.raw);
            }
        } else {
#pragma unroll
            for (int p = 0; p < PX; ++p)
                if (ow0 + p < OW) o[p] = from_f32<TO>(v[p]);
        }
    }
}

template <int PX>
__global__ void __launch_bounds__(256)
upsample_argmax_kernel(const float* __restrict__ x, int IH, int IW, int C, uint8_t* __restrict__ mask, int OH, int OW,
                       float sh, float sw, const void* __restrict__ labels, int label_dtype, int ignore_label,
                       unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned int s_hist[];  // [C*C] when hist != nullptr
    if (hist) {
        for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_hist[i] = 0u;
        __syncthreads();
    }
    const int groups_w = (OW + PX - 1) / PX;
    const int gw = blockIdx.x * blockDim.x + threadIdx.x;
    const int oh = blockIdx.y, n = blockIdx.z;
    if (gw < groups_w) {
        int y0, y1;
        float wy;
        cab_bilinear_tap(oh, sh, IH, y0, y1, wy);
        const float* r0 = x + (static_cast<long long>(n) * IH + y0) * IW * C;
        const float* r1 = x + (static_cast<long long>(n) * IH + y1) * IW * C;
        const int ow0 = gw * PX;
        float best[PX];
        int arg[PX];
        int x0[PX], x1[PX];
        float wx[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            cab_bilinear_tap(min(ow0 + p, OW - 1), sw, IW, x0[p], x1[p], wx[p]);
            best[p] = -INFINITY;
            arg[p] = 0;
        }
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const float a = (1.f - wx[p]) * __ldg(r0 + x0[p] * C + c) + wx[p] * __ldg(r0 + x1[p] * C + c);
                const float b = (1.f - wx[p]) * __ldg(r1 + x0[p] * C + c) + wx[p] * __ldg(r1 + x1[p] * C + c);
                const float v = (1.f - wy) * a + wy * b;
                if (v > best[p]) {  // strict: first maximum wins (torch.argmax)
                    best[p] = v;
                    arg[p] = c;
                }
            }
        }
        const long long obase = (static_cast<long long>(n) * OH + oh) * OW + ow0;
        if (mask) {
            if (ow0 + PX <= OW && (OW % PX) == 0 && PX == 8) {
                uint32_t lo = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
                uint32_t hi = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
                *reinterpret_cast<uint2*>(mask + obase) = make_uint2(lo, hi);
            } else {
#pragma unroll
                for (int p = 0; p < PX; ++p)
                    if (ow0 + p < OW) mask[obase + p] = static_cast<uint8_t>(arg[p]);
            }
        }
        if (hist) {
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                if (ow0 + p >= OW) continue;
                long long lb = label_dtype == 0 ? reinterpret_cast<const long long*>(labels)[obase + p]
                                                : static_cast<long long>(reinterpret_cast<const uint8_t*>(labels)[obase + p]);
                if (lb == ignore_label) continue;
                const int l = static_cast<int>(min(max(lb, 0LL), static_cast<long long>(C - 1)));
                atomicAdd(&s_hist[arg[p] * C + l], 1u);
            }
        }
    }
    if (hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < C * C; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&hist[i], static_cast<unsigned long long>(s_hist[i]));
    }
}

template <typename TP, typename TL>
__global__ void __launch_bounds__(256)
confusion_hist_kernel(const TP* __restrict__ pred, const TL* __restrict__ labels, long long n, int C, int ignore_label,
                      unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned int s_hist[];
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long lb = static_cast<long long>(labels[i]);
        if (lb == ignore_label) continue;
        const long long pr = static_cast<long long>(pred[i]);
        const int l = static_cast<int>(min(max(lb, 0LL), static_cast<long long>(C - 1)));
        const int p = static_cast<int>(min(max(pr, 0LL), static_cast<long long>(C - 1)));
        atomicAdd(&s_hist[p * C + l], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], static_cast<unsigned long long>(s_hist[i]));
}

}  // namespace

extern "C" int cabinet_bilinear_nhwc(const void* x, long long ldx, int x_dtype, void* y, long long ldy, int y_dtype,
                                     int N, int IH, int IW, int C, int OH, int OW, cabinet_stream_t stream) {
    CAB_REQUIRE(x && y && IH > 0 && IW > 0 && OH > 0 && OW > 0 && C > 0 && ldx >= C && ldy >= C,
                "bilinear_nhwc: bad arguments");
    if (N == 0) return CABINET_OK;
    const long long total = static_cast<long long>(N) * OH * OW * ((C + 3) / 4);
    dim3 grid(static_cast<unsigned>(cab_ceil_div(total, 256)));
    const float sh = static_cast<float>(IH) / static_cast<float>(OH), sw = static_cast<float>(IW) / static_cast<float>(OW);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CAB_BL(TI, TO)                                                                                       \
    bilinear_nhwc_kernel<TI, TO><<<grid, 256, 0, s>>>(reinterpret_cast<const TI*>(x), ldx,                   \
                                                      reinterpret_cast<TO*>(y), ldy, IH, IW, C, OH, OW, total, sh, sw)
    if (x_dtype == CABINET_F32 && y_dtype == CABINET_F32) CAB_BL(float, float);
    else if (x_dtype == CABINET_F32) CAB_BL(float, bf16);
    else if (y_dtype == CABINET_F32) CAB_BL(bf16, float);
    else CAB_BL(bf16, bf16);
#undef CAB_BL
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}

extern "C" int cabinet_upsample_logits_nchw(const float* x, int N, int IH, int IW, int C, void* y, int y_dtype, int OH,
                                            int OW, cabinet_stream_t stream) {
    CAB_REQUIRE(x && y && IH > 0 && IW > 0 && OH > 0 && OW > 0 && C > 0, "upsample_logits_nchw: bad arguments");
    CAB_REQUIRE(OH <= 65535 && N <= 65535, "upsample_logits_nchw: OH/N exceed grid limits");
    if (N == 0) return CABINET_OK;
    const float sh = static_cast<float>(IH) / static_cast<float>(OH), sw = static_cast<float>(IW) / static_cast<float>(OW);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (y_dtype == CABINET_F32) {
        constexpr int PX = 4;
        const int groups = (OW + PX - 1) / PX;
        dim3 grid((groups + 127) / 128, OH, N);
        upsample_logits_nchw_kernel<float, PX><<<grid, 128, 0, s>>>(x, IH, IW, C, reinterpret_cast<float*>(y), OH, OW,
                                                                    sh, sw);
    } else {
        constexpr int PX = 8;
        const int groups = (OW + PX - 1) / PX;
        dim3 grid((groups + 127) / 128, OH, N);
        upsample_logits_nchw_kernel<bf16, PX><<<grid, 128, 0, s>>>(x, IH, IW, C, reinterpret_cast<bf16*>(y), OH, OW, sh,
                                                                   sw);
    }
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}

extern "C" int cabinet_upsample_argmax(const float* x, int N, int IH, int IW, int C, uint8_t* mask, int OH, int OW,
                                       const void* labels, int label_dtype, int ignore_label, long long* hist,
                                       cabinet_stream_t stream) {
    CAB_REQUIRE(x && (mask || hist) && IH > 0 && IW > 0 && OH > 0 && OW > 0 && C > 0 && C <= 255,
                "upsample_argmax: bad arguments (1 <= C <= 255)");
    CAB_REQUIRE(!hist || labels, "upsample_argmax: hist requested without labels");
    CAB_REQUIRE(C * C * sizeof(unsigned) <= 48 * 1024, "upsample_argmax: C*C histogram does not fit shared memory");
    CAB_REQUIRE(OH <= 65535 && N <= 65535, "upsample_argmax: OH/N exceed grid limits");
    if (N == 0) return CABINET_OK;
    const float sh = static_cast<float>(IH) / static_cast<float>(OH), sw = static_cast<float>(IW) / static_cast<float>(OW);
    constexpr int PX = 8;
    const int groups = (OW + PX - 1) / PX;
    dim3 grid((groups + 127) / 128, OH, N);
    const size_t smem = hist ? sizeof(unsigned) * C * C : 0;
    upsample_argmax_kernel<PX><<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(
        x, IH, IW, C, mask, OH, OW, sh, sw, labels, label_dtype, ignore_label,
        reinterpret_cast<unsigned long long*>(hist));
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}

extern "C" int cabinet_confusion_hist(const void* pred, int pred_dtype, const void* labels, int label_dtype,
                                      long long n_pixels, int C, int ignore_label, long long* hist,
                                      cabinet_stream_t stream) {
    CAB_REQUIRE(pred && labels && hist && C > 0 && C * C * sizeof(unsigned) <= 48 * 1024,
                "confusion_hist: bad arguments");
    if (n_pixels == 0) return CABINET_OK;
    dim3 grid(static_cast<unsigned>(std::min<long long>(cab_ceil_div(n_pixels, 256 * 16), 148 * 8)));
    grid.x = grid.x ? grid.x : 1;
    const size_t smem = sizeof(unsigned) * C * C;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto* h = reinterpret_cast<unsigned long long*>(hist);
#define CAB_CH(TP, TL)                                                                                             \
    confusion_hist_kernel<TP, TL><<<grid, 256, smem, s>>>(reinterpret_cast<const TP*>(pred),                       \
                                                          reinterpret_cast<const TL*>(labels), n_pixels, C, ignore_label, h)
    if (pred_dtype == 0 && label_dtype == 0) CAB_CH(long long, long long);
    else if (pred_dtype == 0) CAB_CH(long long, uint8_t);
    else if (label_dtype == 0) CAB_CH(uint8_t, long long);
    else CAB_CH(uint8_t, uint8_t);
#undef CAB_CH
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}
