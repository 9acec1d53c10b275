// Fused MobileNetV3 inverted-residual block, channel-major ("transposed") formulation
// (reference: src/models/mobilenetv3.py:102-159), bf16 NHWC:
//
//     expand 1x1 (+BN +act)  ->  depthwise k x k (+BN [+act])  ->  [ project 1x1 (+BN) (+identity) ]
//
// cabinet_mbconv_fused (mbconv_fused.cu) computes the expand GEMM pixel-major (TMEM lane = pixel), stages the expanded
// tile as bf16 in shared memory and runs the depthwise conv from there: per output it pays shared-memory loads, bf16
// unpacks and two CTA-wide barriers per chunk -- it is bound by CUDA-core instruction issue (ncu: 6.0 k warp
// instructions per 64-pixel tile at 50 % issue utilisation, profiles/r02_ncu_mbconv_fused_*.csv).  This kernel turns
// the GEMM around:
//
//   per output tile (TH x TW pixels of one image) and chunk of 128 expanded channels
//     1. MMA1 (tcgen05):  D1[channel][input pixel incl. halo] = W1_chunk[128 ch][Cin + 2] * A1[pixels][Cin + 2]^T
//        TMEM LANE = expanded channel, TMEM COLUMN = pixel of the input patch (row-major, IWT columns per patch row).
//        A1 is the same 4-D TMA box as before (one per 64-channel K block); the expand bias rides in two spare K slots
//        (the MMA warp writes 1.0 there for IN-IMAGE pixels only, so pixels outside the image come out as exactly 0 =
//        the zero padding the depthwise conv sees -- no masking anywhere else).
//     2. depthwise (CUDA cores): thread = (channel, output row segment).  Its input rows are CONSECUTIVE TMEM COLUMNS of
//        its own lane: tcgen05.ld -> fp32 registers -> activation -> FMAs (packed FFMA2 over adjacent output pixels
//        where the register pairs line up).  No shared memory, no bf16 round trip of the expanded tensor, no unpack.
//     3a. PROJECT: the thread's outputs (one channel, NC consecutive pixels) are 2 x 16 contiguous bytes of the
//         M-major (pixel-contiguous) SWIZZLE_128B A operand of MMA2: D2[pixel][Cout] += A2^T W2_chunk^T, accumulated
//         over the chunks in TMEM; after the last chunk TMEM -> +bias (+identity) -> bf16 -> global.
//     3b. !PROJECT (squeeze-excite blocks): outputs go straight to global memory (a warp writes 64 contiguous bytes per
//         pixel) and the per-(image, channel) pooling sums are added as 64-bit fixed-point integers (deterministic).
//
// One persistent CTA per SM: 16 compute warps (4 TMEM lane quarters x 4 pixel segments), one TMA warp, one MMA warp.
// The input patches travel through a ring of up to 4 shared-memory stages (the TMA warp runs whole tiles ahead and
// writes the bias slots of a landed patch), D1 through a ring of 2-3 TMEM stages, A2 is double buffered: TMA latency,
// MMA1 of the next chunks and MMA2 of the previous one all run under the depthwise phase; the warps only meet on
// mbarriers (no CTA-wide barrier in the steady state).  Expanded widths <= 64 are replicated
// twice along the TMEM lanes (rows 64..127 of W1 repeat rows 0..63) so that all 128 lanes have a channel to work on.
#include "tc_common.cuh"

namespace {

constexpr int NCW = 16;                       // compute warps
constexpr int TMA_WARP = NCW, MMA_WARP = NCW + 1;
constexpr int NTHREADS = (NCW + 2) * 32;
constexpr int A2_BUF = 2 * 128 * 128;         // [2 pixel blocks of 64][128 channel rows][128 B]
constexpr int W1_KB = 128 * 128;              // one 64-channel K block of a 128-row expand-weight chunk
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 4;                 // A1 ring (shared memory) / D1 ring (TMEM) depth

struct MtParams {
    int H, W, OH, OW, Cin, Cexp, Cout, cout_pad;
    int nc, tiles_w, tiles_h, num_tiles;
    int act_e, act_dw, has_res, ksteps1, kb, resident;
    int a1_kb_bytes, w1_buf_bytes, w2_buf_bytes, off_w1, off_w2, off_a2, off_b2;
    int ns, a1_stage_bytes;   // A1 ring: stages, bytes per stage (= kb K blocks)
    int ncols, nd, d2col;     // D1 ring in TMEM: columns per stage (= stride), stages; first column of D2
    int nd2, d2_stride;       // D2 buffers (1 | 2) and their column stride
    int nw2;                  // W2 chunk buffers (2, or 1 when shared memory is short)
    int step[3];
    const float* aux;   // [nc][K*K + 1][128] fp32: depthwise taps, depthwise bias (row layout = TMEM lanes)
    const float* b2;
    const bf16* res;
    long long ldres;
    bf16* y;
    long long ldy;
    long long* gap;
    const float* scale;  // project mode: [N][Cexp] squeeze-excite gate applied to the BN output before act_dw, or nullptr
};

struct TileIter {
    int tw, th, n;
    __device__ __forceinline__ void init(int tile, int tiles_w, int tiles_h) {
        tw = tile % tiles_w;
        const int t = tile / tiles_w;
        th = t % tiles_h;
        n = t / tiles_h;
    }
    __device__ __forceinline__ void next(const int* step, int tiles_w, int tiles_h) {
        tw += step[0];
        if (tw >= tiles_w) { tw -= tiles_w; ++th; }
        th += step[1];
        if (th >= tiles_h) { th -= tiles_h; ++n; }
        n += step[2];
    }
};

__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t v0, uint32_t v1) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v0), "r"(v1) : "memory");
}
__device__ __forceinline__ void sts32f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds128f(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// MN-major SWIZZLE_128B operand (64 MN elements x 8 K rows per 1024-byte atom): LBO = bytes between 64-element MN blocks
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// tcgen05.ld 32x32b: N consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tld1(uint32_t a, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(r[0]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tld2(uint32_t a, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tld4(uint32_t a, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3])
                 : "r"(a)
                 : "memory");
}
__device__ __forceinline__ void tld8(uint32_t a, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "r"(a)
                 : "memory");
}
__device__ __forceinline__ void tld16(uint32_t a, float* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(a)
        : "memory");
}
template <int N> __device__ __forceinline__ void tld_row(uint32_t a, float* r) {
    if constexpr (N >= 16) {
        tld16(a, r);
        tld_row<N - 16>(a + 16, r + 16);
    } else if constexpr (N >= 8) {
        tld8(a, r);
        tld_row<N - 8>(a + 8, r + 8);
    } else if constexpr (N >= 4) {
        tld4(a, r);
        tld_row<N - 4>(a + 4, r + 4);
    } else if constexpr (N >= 2) {
        tld2(a, r);
        tld_row<N - 2>(a + 2, r + 2);
    } else if constexpr (N == 1) {
        tld1(a, r);
    }
}

// The block's activations are none / ReLU / hard-swish only (mobilenetv3.py:128-143): one uniform branch, straight-line code
template <int N> __device__ __forceinline__ void act3(float* v, int act) {
    if (act == CABINET_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (act == CABINET_ACT_HSWISH) {  // x * relu6(x + 3) / 6 = x * saturate(x / 6 + 0.5)
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = v[i] * __saturatef(fmaf(v[i], 1.f / 6.f, 0.5f));
    }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi, bool relu) {
    uint32_t d;
    if (relu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

template <int K, int S, int REP, bool PROJECT, int TW>
__global__ void __launch_bounds__(NTHREADS, 1)
mbconv_t_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const MtParams p) {
    constexpr int PAD = (K - 1) / 2;
    constexpr int TH = S == 1 ? 8 : 4;
    constexpr int NR = (S == 1 && REP == 1) ? 2 : 1;                   // output rows per thread
    constexpr int NC = S == 1 ? TW : (REP == 1 ? 8 : 4);               // output columns per thread
    static_assert(TW == 8 || (TW == 16 && S == 1), "tiles: 8 x 16 or 8 x 8 (stride 1), 4 x 8 (stride 2)");
    constexpr int IWT = (TW - 1) * S + K, IHT = (TH - 1) * S + K, NPIX = IWT * IHT;
    constexpr int NIN = (NC - 1) * S + K, NROWS = (NR - 1) * S + K;    // input window of one thread
    constexpr int CH = 128 / REP;                                      // distinct channels per chunk
    constexpr int SPR = TW / NC;                                       // segments per output row
    constexpr int KK = K * K;
    static_assert((TH / NR) * SPR == 4 * REP, "16 compute warps = 4 lane quarters x 4 segments (x replicas)");

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a1_full[MAX_STAGES], a1_ready[MAX_STAGES], a1_free[MAX_STAGES], w1_full[2], w1_free[2],
        w2_full[2], w2_free[2], d1_full[MAX_STAGES], d1_free[MAX_STAGES], a2_full[2], a2_free[2], d2_full[2], d2_free[2];
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sA1 = tc::smem_u32(smem);
    const uint32_t sW1 = sA1 + p.off_w1, sW2 = sA1 + p.off_w2, sA2 = sA1 + p.off_a2, s_b2 = sA1 + p.off_b2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nc = p.nc;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmX);
        tc::prefetch_tmap(&tmW1);
        if (PROJECT) tc::prefetch_tmap(&tmW2);
        for (int i = 0; i < MAX_STAGES; ++i) {
            tc::mbar_init(&a1_full[i], 1);
            tc::mbar_init(&a1_ready[i], 1);
            tc::mbar_init(&a1_free[i], 1);
            tc::mbar_init(&d1_full[i], 1);
            tc::mbar_init(&d1_free[i], NCW);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&w1_full[i], 1);
            tc::mbar_init(&w1_free[i], 1);
            tc::mbar_init(&w2_full[i], 1);
            tc::mbar_init(&w2_free[i], 1);
            tc::mbar_init(&a2_full[i], NCW);
            tc::mbar_init(&d2_full[i], 1);
            tc::mbar_init(&d2_free[i], NCW);
            tc::mbar_init(&a2_free[i], 1);
        }
        tc::mbar_fence_init();
        tc::fence_proxy_async();
    }
    if (warp == MMA_WARP) tc::tmem_alloc(&tmem_base_smem, TMEM_COLS);
    if (PROJECT)
        for (int i = threadIdx.x; i < p.cout_pad; i += NTHREADS) sts32f(s_b2 + 4 * i, i < p.Cout ? __ldg(p.b2 + i) : 0.f);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_smem;
    const int n_my = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int G = n_my * nc;

    if (warp == TMA_WARP) {
        // ============================ TMA producer (+ bias slots of the landed patches) ============================
        TileIter lt, ft;  // next tile to load / to finalize
        lt.init(blockIdx.x, p.tiles_w, p.tiles_h);
        ft.init(blockIdx.x, p.tiles_w, p.tiles_h);
        int issued = 0, finalized = 0;
        auto issue_a1 = [&]() {  // patch of tile `issued` -> stage issued % ns
            const int t = issued, s = t % p.ns;
            if (t >= p.ns) tc::mbar_wait(&a1_free[s], ((t / p.ns) - 1) & 1);
            if (lane == 0) {
                uint8_t* dst = smem + s * p.a1_stage_bytes;
                tc::mbar_expect_tx(&a1_full[s], NPIX * 128 * p.kb);
                for (int kb = 0; kb < p.kb; ++kb)
                    tc::tma_load_4d(dst + kb * p.a1_kb_bytes, &tmX, &a1_full[s], kb * 64, lt.tw * TW * S - PAD,
                                    lt.th * TH * S - PAD, lt.n);
            }
            lt.next(p.step, p.tiles_w, p.tiles_h);
            ++issued;
            __syncwarp();
        };
        auto finalize_a1 = [&]() {
            // bias slots: K columns Cin, Cin + 1 of every staged IN-IMAGE pixel <- 1.0.  Rows outside the image stay all
            // zero (TMA fill), so their expanded value is act(0) = 0: the depthwise conv's zero padding.
            const int t = finalized, s = t % p.ns;
            tc::mbar_wait(&a1_full[s], (t / p.ns) & 1);
            const uint32_t ones = 0x3F803F80u;
            const uint32_t s_ones = sA1 + s * p.a1_stage_bytes + (p.Cin >> 6) * p.a1_kb_bytes;
            const int cq = (p.Cin & 63) >> 3;
            const int ih0 = ft.th * TH * S - PAD, iw0 = ft.tw * TW * S - PAD;
            for (int r = lane; r < NPIX; r += 32) {
                const int iy = r / IWT, ix = r - iy * IWT;
                const int ih = ih0 + iy, iw = iw0 + ix;
                if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) sts32(s_ones + r * 128 + ((cq ^ (r & 7)) << 4), ones);
            }
            ft.next(p.step, p.tiles_w, p.tiles_h);
            ++finalized;
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&a1_ready[s]);
        };
        auto load_w1 = [&](int c, int buf) {
            if (lane == 0) {
                uint8_t* dst = smem + p.off_w1 + buf * p.w1_buf_bytes;
                tc::mbar_expect_tx(&w1_full[buf], p.kb * W1_KB);
                for (int kb = 0; kb < p.kb; ++kb) tc::tma_load_2d(dst + kb * W1_KB, &tmW1, &w1_full[buf], kb * 64, c * 128);
            }
        };
        auto load_w2 = [&](int c, int buf) {
            if (lane == 0) {
                uint8_t* dst = smem + p.off_w2 + buf * p.w2_buf_bytes;
                tc::mbar_expect_tx(&w2_full[buf], p.w2_buf_bytes);
                for (int j = 0; j < CH / 64; ++j)
                    tc::tma_load_2d(dst + j * p.cout_pad * 128, &tmW2, &w2_full[buf], c * CH + j * 64, 0);
            }
        };
        while (issued < n_my && issued < p.ns) issue_a1();
        if (p.resident) {
            for (int c = 0; c < nc; ++c) {
                load_w1(c, c);
                if (PROJECT) load_w2(c, c);
            }
        }
        for (int e = 0; e <= G; ++e) {
            if (e < G) {
                const int it = e / nc, c = e - it * nc;
                // the first chunk of tile it: every MMA1 of tile it - 1 has been issued -> refill its stage
                if (c == 0 && it >= 1 && issued < n_my) issue_a1();
                // patches needed next: the tile of chunk e + 1 (MMA1 runs ahead of the depthwise phase)
                const int target = min(e + 1, G - 1) / nc;
                while (finalized <= target && finalized < issued) finalize_a1();
                if (!p.resident) {
                    if (e >= 2) tc::mbar_wait(&w1_free[e & 1], ((e >> 1) - 1) & 1);
                    load_w1(c, e & 1);
                }
            }
            if (PROJECT && !p.resident && e >= 1) {  // W2 of chunk e - 1: one step behind W1, it is needed a phase later
                const int g = e - 1, wb = g % p.nw2;
                if (g >= p.nw2) tc::mbar_wait(&w2_free[wb], ((g / p.nw2) - 1) & 1);
                load_w2(g % nc, wb);
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        // ============================ MMA issue ============================
        const uint32_t leader = tc::elect_one();
        const uint32_t idesc1 = tc::make_idesc_bf16(128, p.ncols);
        const uint32_t idesc2 = tc::make_idesc_bf16(128, p.cout_pad) | (1u << 15);  // A (= A2) is MN-major
        auto mma1 = [&](int g) {
            const int it = g / nc, c = g - it * nc;
            const int s = it % p.ns, ds = g % p.nd;
            const int wbuf = p.resident ? c : (g & 1);
            if (c == 0) tc::mbar_wait(&a1_ready[s], (it / p.ns) & 1);
            tc::mbar_wait(&w1_full[wbuf], p.resident ? 0 : ((g >> 1) & 1));
            if (g >= p.nd) tc::mbar_wait(&d1_free[ds], ((g / p.nd) - 1) & 1);
            tc::tc_fence_after();
            const uint64_t a_desc = tc::make_desc_sw128(sW1 + wbuf * p.w1_buf_bytes);
            const uint64_t b_desc = tc::make_desc_sw128(sA1 + s * p.a1_stage_bytes);
            const uint32_t d = tmem + ds * p.ncols;
            for (int ks = 0; ks < p.ksteps1; ++ks) {
                const uint64_t ko = static_cast<uint64_t>((ks & 3) * 2);
                tc::umma_bf16_if(leader, d, a_desc + static_cast<uint64_t>(((ks >> 2) * W1_KB) >> 4) + ko,
                                 b_desc + static_cast<uint64_t>(((ks >> 2) * p.a1_kb_bytes) >> 4) + ko, idesc1, ks > 0 ? 1u : 0u);
            }
            tc::umma_commit_if(leader, &d1_full[ds]);
            if (!p.resident) tc::umma_commit_if(leader, &w1_free[g & 1]);
            if (c == nc - 1) tc::umma_commit_if(leader, &a1_free[s]);
        };
        for (int j = 0; j < p.nd - 1 && j < G; ++j) mma1(j);  // MMA1 runs nd - 1 chunks ahead of the depthwise phase
        for (int g = 0; g < G; ++g) {
            if (g + p.nd - 1 < G) mma1(g + p.nd - 1);
            if (PROJECT) {
                const int it = g / nc, c = g - it * nc;
                const int buf = g & 1, wbuf = p.resident ? c : g % p.nw2;
                tc::mbar_wait(&w2_full[wbuf], p.resident ? 0 : ((g / p.nw2) & 1));
                tc::mbar_wait(&a2_full[buf], (g >> 1) & 1);
                const int db = it % p.nd2;
                if (c == 0 && it >= p.nd2) tc::mbar_wait(&d2_free[db], ((it / p.nd2) - 1) & 1);
                tc::tc_fence_after();
                const uint64_t a_desc = make_desc_mn_sw128(sA2 + buf * A2_BUF, 128 * 128);
                const uint64_t b_desc = tc::make_desc_sw128(sW2 + wbuf * p.w2_buf_bytes);
#pragma unroll
                for (int ks = 0; ks < CH / 16; ++ks)
                    tc::umma_bf16_if(leader, tmem + p.d2col + db * p.d2_stride, a_desc + static_cast<uint64_t>(ks * (2048 >> 4)),
                                     b_desc + static_cast<uint64_t>(((ks >> 2) * p.cout_pad * 128) >> 4) +
                                         static_cast<uint64_t>((ks & 3) * 2),
                                     idesc2, (c > 0 || ks > 0) ? 1u : 0u);
                tc::umma_commit_if(leader, &a2_free[buf]);
                if (!p.resident) tc::umma_commit_if(leader, &w2_free[wbuf]);
                if (c == nc - 1) tc::umma_commit_if(leader, &d2_full[db]);
            }
        }
        __syncwarp();
    } else {
        // ============================ compute warps ============================
        const int q = warp & 3, grp = warp >> 2;
        const int l = q * 32 + lane;                 // TMEM lane
        const int kch = l % CH;                      // channel inside the chunk
        const int ps = (l / CH) * 4 + grp;           // pixel segment
        const int oy0 = (ps / SPR) * NR, ox0 = (ps % SPR) * NC;
        const uint32_t tlane = tmem + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t win0 = static_cast<uint32_t>(oy0 * S * IWT + ox0 * S);  // first TMEM column of this thread's input window
        const bool relu_dw = p.act_dw == CABINET_ACT_RELU;
        TileIter ti;
        ti.init(blockIdx.x, p.tiles_w, p.tiles_h);
        // ---- epilogue 2 of one tile: D2[pixel lane][Cout] -> +bias (+identity) -> bf16 -> global.  Deferred by one chunk: it
        // runs after the first depthwise pass of the NEXT tile, so the compute warps never sit out the a2_full -> MMA2 ->
        // d2_full round trip (MMA2 of the next tile waits for d2_free instead, on the MMA warp).
        auto epilogue2 = [&](int e_it, int e_oh0, int e_ow0, int e_n) {
            const int m = l;
            const int oh = e_oh0 + m / TW, ow = e_ow0 + m % TW;
            const bool valid = m < TH * TW && oh < p.OH && ow < p.OW;
            const long long pix = (static_cast<long long>(e_n) * p.OH + oh) * p.OW + ow;
            const int db = e_it % p.nd2;
            tc::mbar_wait(&d2_full[db], (e_it / p.nd2) & 1);
            tc::tc_fence_after();
            for (int j16 = grp; j16 * 16 < p.cout_pad; j16 += 4) {
                uint32_t v[16];
                tc::tmem_ld16(tlane + p.d2col + db * p.d2_stride + j16 * 16, v);
                tc::tmem_ld_wait();
                const int co0 = j16 * 16;
                float f[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bb = lds128f(s_b2 + (co0 + 4 * j) * 4);
                    f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bb.x;
                    f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bb.y;
                    f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bb.z;
                    f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bb.w;
                }
                if (valid && co0 < p.Cout) {
                    const bool second = co0 + 8 < p.Cout;
                    if (p.has_res) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.ldres + co0);
                        Vec16<bf16> r0, r1;
                        r0.raw = __ldg(rp);
                        r1.raw = second ? __ldg(rp + 1) : make_uint4(0, 0, 0, 0);
                        float rf[16];
                        r0.unpack(rf);
                        r1.unpack(rf + 8);
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] += rf[j];
                    }
                    Vec16<bf16> o0, o1;
                    o0.pack(f);
                    o1.pack(f + 8);
                    uint4* yp = reinterpret_cast<uint4*>(p.y + pix * p.ldy + co0);
                    yp[0] = o0.raw;
                    if (second) yp[1] = o1.raw;
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&d2_free[db]);
        };
        int pend_it = -1, pend_oh0 = 0, pend_ow0 = 0, pend_n = 0;
        int g = 0, ds = 0, dphase = 0;  // D1 ring position of chunk g
        float tap[KK], bias = 0.f;
        int tap_chunk = -1;
        for (int it = 0; it < n_my; ++it, ti.next(p.step, p.tiles_w, p.tiles_h)) {
            const int ow0 = ti.tw * TW, oh0 = ti.th * TH, n = ti.n;
            for (int c = 0; c < nc; ++c, ++g) {
                const int buf = g & 1;
                // this thread's channel constants (coalesced over the lanes; L1 / L2 resident); single-chunk blocks load
                // them once for the whole kernel
                if (S == 1 || c != tap_chunk) {  // (kept across tiles only in the small stride-2 kernels: registers)
                    const float* ax = p.aux + static_cast<size_t>(c) * (KK + 1) * 128 + l;
#pragma unroll
                    for (int t = 0; t < KK; ++t) tap[t] = __ldg(ax + t * 128);
                    bias = __ldg(ax + KK * 128);
                    tap_chunk = c;
                }
                float acc[NR][NC];
#pragma unroll
                for (int a = 0; a < NR; ++a)
#pragma unroll
                    for (int j = 0; j < NC; ++j) acc[a][j] = bias;

                tc::mbar_wait(&d1_full[ds], dphase);
                tc::tc_fence_after();
                const uint32_t t0 = tlane + ds * p.ncols + win0;
                float in[2][NIN];
                tld_row<NIN>(t0, in[0]);
#pragma unroll
                for (int r = 0; r < NROWS; ++r) {
                    tc::tmem_ld_wait();
                    if (r + 1 < NROWS) tld_row<NIN>(t0 + (r + 1) * IWT, in[(r + 1) & 1]);
                    if (r == NROWS - 1) {  // all of this chunk's D1 reads have landed: MMA1 of chunk g + 2 may overwrite it
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&d1_free[ds]);
                        if (++ds == p.nd) { ds = 0; dphase ^= 1; }
                    }
                    float* v = in[r & 1];
                    act3<NIN>(v, p.act_e);
#pragma unroll
                    for (int a = 0; a < NR; ++a) {
                        const int ky = r - a * S;  // compile-time after unrolling
                        if (ky < 0 || ky >= K) continue;
#pragma unroll
                        for (int kx = 0; kx < K; ++kx) {
                            const float w = tap[ky * K + kx];
                            if (S == 1 && (kx & 1) == 0) {  // (v[j + kx], v[j + kx + 1]) is an aligned register pair
                                const float2 w2 = make_float2(w, w);
#pragma unroll
                                for (int j = 0; j < NC; j += 2) {
                                    float2 a2 = make_float2(acc[a][j], acc[a][j + 1]);
                                    cab_ffma2(a2, make_float2(v[j + kx], v[j + kx + 1]), w2);
                                    acc[a][j] = a2.x;
                                    acc[a][j + 1] = a2.y;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < NC; ++j) acc[a][j] = fmaf(v[j * S + kx], w, acc[a][j]);
                            }
                        }
                    }
                }
                const int cg = c * CH + kch;  // expanded channel of this thread
                if (!PROJECT) {
                    // ---- pooling sums (the BN output BEFORE act_dw, mobilenetv3.py:137-143) + direct global store
                    float gsum = 0.f;
                    const bool ch_ok = cg < p.Cexp;
                    const int ldy = static_cast<int>(p.ldy);
                    if (oh0 + TH <= p.OH && ow0 + TW <= p.OW) {  // interior tile: no per-pixel bounds tests
#pragma unroll
                        for (int a = 0; a < NR; ++a) {
                            bf16* yrow = p.y + ((static_cast<long long>(n) * p.OH + oh0 + oy0 + a) * p.OW + ow0 + ox0) * p.ldy + cg;
#pragma unroll
                            for (int j = 0; j < NC; ++j) gsum += acc[a][j];
                            if (!relu_dw) act3<NC>(acc[a], p.act_dw);
                            if (ch_ok) {
#pragma unroll
                                for (int j = 0; j < NC; j += 2) {
                                    const uint32_t h = pack_bf16x2(acc[a][j], acc[a][j + 1], relu_dw);
                                    *reinterpret_cast<unsigned short*>(yrow + j * ldy) = static_cast<unsigned short>(h & 0xffffu);
                                    *reinterpret_cast<unsigned short*>(yrow + (j + 1) * ldy) = static_cast<unsigned short>(h >> 16);
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int a = 0; a < NR; ++a) {
                            const int oh = oh0 + oy0 + a;
                            bf16* yrow = p.y + ((static_cast<long long>(n) * p.OH + oh) * p.OW + ow0 + ox0) * p.ldy + cg;
#pragma unroll
                            for (int j = 0; j < NC; ++j)
                                if (oh < p.OH && ow0 + ox0 + j < p.OW) gsum += acc[a][j];
                            if (!relu_dw) act3<NC>(acc[a], p.act_dw);
                            if (ch_ok && oh < p.OH) {
#pragma unroll
                                for (int j = 0; j < NC; ++j) {
                                    if (ow0 + ox0 + j < p.OW) {
                                        const float o = relu_dw ? fmaxf(acc[a][j], 0.f) : acc[a][j];
                                        yrow[j * ldy] = __float2bfloat16_rn(o);
                                    }
                                }
                            }
                        }
                    }
                    if (p.gap && ch_ok)
                        atomicAdd(reinterpret_cast<unsigned long long*>(p.gap + static_cast<long long>(n) * p.Cexp + cg),
                                  static_cast<unsigned long long>(__float2ll_rn(gsum * CABINET_GAP_FIXED_ONE)));
                } else {
                    // ---- A2 (M-major SWIZZLE_128B): row = channel, 128 B = 64 consecutive output pixels
                    if (p.scale) {  // squeeze-excite: act(gate * BN output), mobilenetv3.py:137-143 (gate >= 0)
                        const float sc = cg < p.Cexp ? __ldg(p.scale + static_cast<long long>(n) * p.Cexp + cg) : 0.f;
#pragma unroll
                        for (int a = 0; a < NR; ++a)
#pragma unroll
                            for (int j = 0; j < NC; ++j) acc[a][j] *= sc;
                    }
                    if (!relu_dw) {
#pragma unroll
                        for (int a = 0; a < NR; ++a) act3<NC>(acc[a], p.act_dw);
                    }
                    if (g >= 2) tc::mbar_wait(&a2_free[buf], ((g >> 1) - 1) & 1);
                    const uint32_t a2row = sA2 + buf * A2_BUF + kch * 128;
#pragma unroll
                    for (int a = 0; a < NR; ++a) {
                        const int m0 = (oy0 + a) * TW + ox0;  // first output pixel of the segment
                        const uint32_t blk = a2row + (m0 >> 6) * (128 * 128);
                        uint32_t h[NC / 2];
#pragma unroll
                        for (int j = 0; j < NC / 2; ++j) h[j] = pack_bf16x2(acc[a][2 * j], acc[a][2 * j + 1], relu_dw);
                        if constexpr (NC >= 8) {
#pragma unroll
                            for (int j8 = 0; j8 < NC / 8; ++j8) {
                                const uint32_t cidx = static_cast<uint32_t>(((m0 & 63) >> 3) + j8);
                                tc::sts128(blk + ((cidx ^ (kch & 7)) << 4),
                                           make_uint4(h[4 * j8], h[4 * j8 + 1], h[4 * j8 + 2], h[4 * j8 + 3]));
                            }
                        } else {
                            const uint32_t cidx = static_cast<uint32_t>((m0 & 63) >> 3);
                            sts64(blk + ((cidx ^ (kch & 7)) << 4) + (m0 & 7) * 2, h[0], h[1]);
                        }
                    }
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&a2_full[buf]);
                    if (c == 0 && pend_it >= 0) {
                        epilogue2(pend_it, pend_oh0, pend_ow0, pend_n);
                        pend_it = -1;
                    }
                }
            }
            if (PROJECT) {
                pend_it = it;
                pend_oh0 = oh0;
                pend_ow0 = ow0;
                pend_n = n;
            }
        }
        if (PROJECT && pend_it >= 0) epilogue2(pend_it, pend_oh0, pend_ow0, pend_n);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem, TMEM_COLS);
    }
}

template <int K, int S, int REP, bool PROJECT, int TW>
int launch_mt(const void* x, long long ldx, int N, const void* w1, const void* w2, MtParams p, cudaStream_t st) {
    constexpr int TH = S == 1 ? 8 : 4;
    constexpr int IWT = (TW - 1) * S + K, IHT = (TH - 1) * S + K, NPIX = IWT * IHT;
    constexpr int CH = 128 / REP;
    p.ncols = (NPIX + 15) / 16 * 16;
    // TMEM: D1 ring from column 0, D2 buffer(s) at the top.  Two D1 stages first, then a second D2 buffer (the epilogue
    // of one tile and MMA2 of the next stop serialising), then more D1 stages.
    p.d2_stride = (p.cout_pad + 31) / 32 * 32;
    p.nd2 = PROJECT && 2 * p.ncols + 2 * p.d2_stride <= TMEM_COLS ? 2 : 1;
    p.d2col = PROJECT ? TMEM_COLS - p.nd2 * p.d2_stride : TMEM_COLS;
    p.nd = std::min(MAX_STAGES, p.d2col / p.ncols);
    if (p.ncols > 256 || p.nd < 2) {
        cabinet_set_error("mbconv_t: TMEM budget (patch %d pixels, cout %d)", NPIX, p.Cout);
        return CABINET_ERR_INVALID;
    }
    p.tiles_w = (p.OW + TW - 1) / TW;
    p.tiles_h = (p.OH + TH - 1) / TH;
    const long long tiles = static_cast<long long>(N) * p.tiles_w * p.tiles_h;
    CAB_REQUIRE(tiles < (1LL << 31), "mbconv_t: too many tiles");
    p.num_tiles = static_cast<int>(tiles);
    p.a1_kb_bytes = ((p.ncols * 128 + 1023) / 1024) * 1024;
    p.w1_buf_bytes = p.kb * W1_KB;
    p.w2_buf_bytes = PROJECT ? (CH / 64) * p.cout_pad * 128 : 0;
    p.resident = p.nc <= 2 ? 1 : 0;
    p.a1_stage_bytes = p.kb * p.a1_kb_bytes;
    // everything but the A1 ring; the ring takes what is left of 226 KB (static shared memory shares the 227 KB)
    p.nw2 = 2;
    int fixed = 2 * p.w1_buf_bytes + 2 * p.w2_buf_bytes + (PROJECT ? 2 * A2_BUF : 0) + 1024 + p.cout_pad * 4 + 1024;
    if (PROJECT && !p.resident && (226 * 1024 - fixed) / p.a1_stage_bytes < 2) {  // W2 single buffered: one more patch stage
        p.nw2 = 1;
        fixed -= p.w2_buf_bytes;
    }
    p.ns = std::min(MAX_STAGES, (226 * 1024 - fixed) / p.a1_stage_bytes);
    if (p.ns < 1) {
        cabinet_set_error("mbconv_t: shared-memory budget (%d bytes + the input patch)", fixed);
        return CABINET_ERR_INVALID;
    }
    p.off_w1 = p.ns * p.a1_stage_bytes;
    p.off_w2 = p.off_w1 + 2 * p.w1_buf_bytes;
    p.off_a2 = ((p.off_w2 + p.nw2 * p.w2_buf_bytes + 1023) / 1024) * 1024;
    p.off_b2 = p.off_a2 + (PROJECT ? 2 * A2_BUF : 0);
    const size_t smem = static_cast<size_t>(p.off_b2) + p.cout_pad * 4 + 1024;
    CUtensorMap tmX, tmW1, tmW2;
    {
        const uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)N};
        const uint64_t strides[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * p.W, (uint64_t)ldx * 2 * p.W * p.H};
        const uint32_t box[4] = {64, (uint32_t)IWT, (uint32_t)IHT, 1};
        int rc = cab_make_tmap_bf16(&tmX, x, 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)p.kb * 64, (uint64_t)p.nc * 128};
        const uint64_t strides[1] = {(uint64_t)p.kb * 128};
        const uint32_t box[2] = {64, 128};
        int rc = cab_make_tmap_bf16(&tmW1, w1, 2, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    if (PROJECT) {
        const uint64_t kk = static_cast<uint64_t>((p.Cexp + 63) / 64) * 64;
        const uint64_t dims[2] = {kk, (uint64_t)p.cout_pad};
        const uint64_t strides[1] = {kk * 2};
        const uint32_t box[2] = {64, (uint32_t)p.cout_pad};
        int rc = cab_make_tmap_bf16(&tmW2, w2, 2, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    } else {
        tmW2 = tmW1;
    }
    int dev = 0, sms = 148;
    CAB_CUDA(cudaGetDevice(&dev));
    CAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = static_cast<int>(std::min<long long>(tiles, sms));
    p.step[0] = grid % p.tiles_w;
    p.step[1] = (grid / p.tiles_w) % p.tiles_h;
    p.step[2] = grid / (p.tiles_w * p.tiles_h);
    static bool attr_done = false;
    if (!attr_done) {
        CAB_CUDA(cudaFuncSetAttribute(mbconv_t_kernel<K, S, REP, PROJECT, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        attr_done = true;
    }
    mbconv_t_kernel<K, S, REP, PROJECT, TW><<<grid, NTHREADS, smem, st>>>(tmX, tmW1, tmW2, p);
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}

}  // namespace

extern "C" int cabinet_mbconv_t(const void* x, long long ldx, int N, int H, int W, int Cin, const void* w_expand_t,
                                const float* aux_t, int Cexp, int act_expand, int k, int stride, int act_dw,
                                const void* w_project, const float* b_project, int Cout, int residual, void* y,
                                long long ldy, int OH, int OW, long long* gap_sum, const float* se_scale,
                                cabinet_stream_t stream) {
    CAB_REQUIRE(x && w_expand_t && aux_t && y, "mbconv_t: null pointer");
    CAB_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2), "mbconv_t: k must be 3|5 and stride 1|2");
    CAB_REQUIRE(N >= 0 && H > 0 && W > 0 && Cin > 0 && Cin <= 248 && Cin % 8 == 0 && Cin % 64 <= 56 && Cexp > 0 &&
                    Cexp % 8 == 0 && Cexp <= 1024,
                "mbconv_t: needs Cin %% 8 == 0, Cin %% 64 <= 56, Cin <= 248 and Cexp %% 8 == 0 (got Cin %d, Cexp %d)", Cin, Cexp);
    const int pad = (k - 1) / 2;
    CAB_REQUIRE(act_expand >= CABINET_ACT_NONE && act_expand <= CABINET_ACT_HSWISH && act_dw >= CABINET_ACT_NONE &&
                    act_dw <= CABINET_ACT_HSWISH,
                "mbconv_t: activations are none / ReLU / hard-swish");
    CAB_REQUIRE(OH == (H + 2 * pad - k) / stride + 1 && OW == (W + 2 * pad - k) / stride + 1, "mbconv_t: inconsistent output size");
    CAB_REQUIRE(ldx % 8 == 0 && ldx >= Cin && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(w_expand_t) & 15) == 0 && (reinterpret_cast<uintptr_t>(aux_t) & 3) == 0,
                "mbconv_t: alignment");
    const bool project = w_project != nullptr;
    if (project) {
        CAB_REQUIRE(b_project && Cout > 0 && Cout <= 128 && Cout % 8 == 0 && ldy >= Cout && ldy % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_project) & 15) == 0,
                    "mbconv_t: project needs a bias, Cout %% 8 == 0, Cout <= 128 and 16-byte aligned output pixels");
        CAB_REQUIRE(!residual || (stride == 1 && Cin == Cout), "mbconv_t: identity needs stride 1, Cin == Cout");
        CAB_REQUIRE(!gap_sum, "mbconv_t: pooling sums exist in the depthwise-output mode only");
    } else {
        CAB_REQUIRE(ldy >= Cexp && !residual && !se_scale, "mbconv_t: depthwise-output mode writes Cexp channels, no identity, no gate");
    }
    if (N == 0) return CABINET_OK;
    MtParams p;
    p.H = H; p.W = W; p.OH = OH; p.OW = OW; p.Cin = Cin; p.Cexp = Cexp; p.Cout = project ? Cout : 0;
    p.cout_pad = project ? (Cout + 15) / 16 * 16 : 0;
    const int rep = Cexp <= 64 ? 2 : 1;
    p.nc = rep == 2 ? 1 : (Cexp + 127) / 128;
    p.act_e = act_expand; p.act_dw = act_dw; p.has_res = residual ? 1 : 0;
    p.ksteps1 = (Cin + 2 + 15) / 16;
    p.kb = Cin / 64 + 1;
    p.aux = aux_t; p.b2 = b_project;
    p.res = reinterpret_cast<const bf16*>(x); p.ldres = ldx;
    p.y = reinterpret_cast<bf16*>(y); p.ldy = ldy; p.gap = gap_sum; p.scale = se_scale;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CAB_MT(K_, S_, R_, TWP_, TWD_)                                                          \
    (project ? launch_mt<K_, S_, R_, true, TWP_>(x, ldx, N, w_expand_t, w_project, p, st)       \
             : launch_mt<K_, S_, R_, false, TWD_>(x, ldx, N, w_expand_t, w_project, p, st))
    if (k == 3 && stride == 1) return rep == 2 ? CAB_MT(3, 1, 2, 16, 16) : CAB_MT(3, 1, 1, 16, 16);
    if (k == 3 && stride == 2) return rep == 2 ? CAB_MT(3, 2, 2, 8, 8) : CAB_MT(3, 2, 1, 8, 8);
    // k = 5: 8 x 16 tiles (12 x 20 patch = 240 TMEM columns) leave no room for D2 -> 8 x 8 tiles in project mode
    if (k == 5 && stride == 1) return rep == 2 ? CAB_MT(5, 1, 2, 8, 16) : CAB_MT(5, 1, 1, 8, 16);
    // k = 5 stride 2: 4 x 8 tiles behind an 11 x 19 patch (224 TMEM columns): depthwise-output mode only
    CAB_REQUIRE(!project, "mbconv_t: k5 stride 2 exists in the depthwise-output mode only (TMEM budget)");
    return rep == 2 ? launch_mt<5, 2, 2, false, 8>(x, ldx, N, w_expand_t, w_project, p, st)
                    : launch_mt<5, 2, 1, false, 8>(x, ldx, N, w_expand_t, w_project, p, st);
#undef CAB_MT
}
