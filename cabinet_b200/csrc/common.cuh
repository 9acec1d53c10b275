// Shared device/host helpers for libcabinet_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>

#include "../../include/cabinet_b200.h"

// ----------------------------------------------------------------------------- error plumbing
void cabinet_set_error(const char* fmt, ...);

#define CAB_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            cabinet_set_error(__VA_ARGS__);    \
            return CABINET_ERR_INVALID;        \
        }                                      \
    } while (0)

#define CAB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            cabinet_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                              __LINE__);                                                        \
            return CABINET_ERR_CUDA;                                                            \
        }                                                                                       \
    } while (0)

#define CAB_LAUNCH_CHECK() CAB_CUDA(cudaGetLastError())

static inline int cab_dtype_size(int dt) { return dt == CABINET_F32 ? 4 : 2; }
static inline long long cab_ceil_div(long long a, long long b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------- device helpers
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float cab_act(float v, int act) {
    // reference: src/models/mobilenetv3.py:38-65 (relu6(x+3)/6, x*hsig(x)); nn.ReLU; nn.Sigmoid
    switch (act) {
        case CABINET_ACT_RELU: return fmaxf(v, 0.f);
        case CABINET_ACT_HSWISH: return v * (fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f);
        case CABINET_ACT_HSIGMOID: return fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f;
        case CABINET_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
        default: return v;
    }
}

// Activation over a small register array: ONE uniform branch per call, straight-line code per case
// (a per-element switch compiles to an indirect branch per element and serialises the whole epilogue).
template <int N> __device__ __forceinline__ void cab_act_vec(float* v, int act) {
    if (act == CABINET_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (act == CABINET_ACT_HSWISH) {  // relu6(x + 3) / 6 == saturate(x / 6 + 0.5): one FFMA.SAT + one FMUL
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = v[i] * __saturatef(fmaf(v[i], 1.f / 6.f, 0.5f));
    } else if (act == CABINET_ACT_HSIGMOID) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __saturatef(fmaf(v[i], 1.f / 6.f, 0.5f));
    } else if (act == CABINET_ACT_SIGMOID) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = 1.f / (1.f + __expf(-v[i]));
    }
}

// Packed fp32 FMA (Blackwell FFMA2): acc.{x,y} += x.{x,y} * w.{x,y} in ONE instruction -- two IEEE fp32 FMAs, bit-identical
// to two fmaf() calls.  The depthwise kernels keep (channel c, channel c+1) pairs in adjacent registers for this.
__device__ __forceinline__ void cab_ffma2(float2& acc, const float2 x, const float2 w) {
    uint64_t a, b, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(x.x), "f"(x.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(w.x), "f"(w.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(b), "l"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a));
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T: 4 floats or 8 bf16
template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    float4 raw;
    __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
    __device__ __forceinline__ void unpack(float* f) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
    __device__ __forceinline__ void pack(const float* f) { raw = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec16<bf16> {
    static constexpr int N = 8;
    uint4 raw;
    __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
    __device__ __forceinline__ void unpack(float* f) const {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // bf16 -> f32 is a 16-bit shift
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ void pack(const float* f) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        raw = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// Bilinear source taps, align_corners=False (ATen area_pixel_compute_source_index; reference call sites
// src/models/cabinet.py:228-245, src/models/cab.py:70-72). scale = in/out as float.
__device__ __forceinline__ void cab_bilinear_tap(int dst, float scale, int in_size, int& i0, int& i1, float& w1) {
    float src = fmaxf((static_cast<float>(dst) + 0.5f) * scale - 0.5f, 0.f);
    i0 = min(static_cast<int>(src), in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    w1 = src - static_cast<float>(i0);
}
