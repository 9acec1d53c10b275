// Fused MobileNetV3 inverted-residual block (reference: src/models/mobilenetv3.py:102-159), bf16 NHWC:
//
//     expand 1x1 (+BN +act)  ->  depthwise k x k (+BN [+act])  ->  [ project 1x1 (+BN) (+identity) ]
//
// The expanded activation is the largest tensor of every block (3x-6x the block input); unfused it is written to
// HBM by the expand GEMM and read back by the depthwise kernel.  Here it only ever exists in shared memory:
//
//   per output tile (TH x TW pixels of one image), per chunk of <= 64 expanded channels
//     1. MMA1 (tcgen05, TMEM):  D1[input pixels incl. halo][64] = A1[pixels][Cin + 2] * W1_chunk^T
//        A1 is one 4-D TMA box {64 ch, IWT, IHT, 1} of the block input per 64-channel K block (OOB zero fill = conv
//        padding of the input, and the zero tail of the last K block); Cin <= 248 -> up to 4 K blocks.
//        The expand bias rides in the GEMM: the control warp writes 1.0 into K slots Cin, Cin+1 of every A1 row and
//        W1 carries (bias_hi, bias_lo) there, so epilogue 1 needs no per-element add.
//     2. epilogue 1 (CUDA cores):  TMEM -> act -> 0 outside the image (the depthwise conv pads the EXPANDED
//        map with zeros) -> bf16 -> E tile in shared memory (144-byte pixel pitch: conflict-free 16-byte stores).
//     3. depthwise (CUDA cores):  warp = output row (segment), lane = channel pair, taps in registers, LDS.32 per
//        lane; result -> A2 tile [out pixels][64] in the 128B-swizzled K-major UMMA layout.
//     4a. PROJECT:  MMA2  D2[out pixels][Cout] += A2 * W2_chunk^T  (accumulates over the chunks in TMEM); after the
//         last chunk: TMEM -> +bias (+ residual) -> bf16 -> staging -> TMA store.
//     4b. !PROJECT (blocks with squeeze-excite): A2 is TMA-stored as the depthwise output of this chunk and the
//         per-(image, channel) sums for the SE pooling are accumulated; scale/act and the project GEMM follow as
//         separate kernels (the SE gate needs the global mean first).
//
// Persistent CTAs, two per SM (<= 113 KB shared memory, 256 TMEM columns each): the phases of one CTA are serial,
// the second CTA fills the gaps.  Warps 0-7 compute, warp 8 issues TMA and MMAs (MMA1 runs one chunk ahead).
#include "tc_common.cuh"

namespace {

constexpr int NCW = 8;                      // compute warps
constexpr int NTHREADS = (NCW + 1) * 32;    // + control warp
constexpr int E_PITCH = 144;                // bytes per staged expanded pixel (64 ch bf16 + 16 B skew)
constexpr int A2_BYTES = 128 * 128;
constexpr int W1_KB = 64 * 128;              // one K block (64 input channels) of a 64-channel expand-weight chunk
constexpr int TMEM_COLS = 256;

struct MbParams {
    int H, W, OH, OW, Cin, Cexp, Cout, cout_pad;
    int n_chunks, resident, tiles_w, tiles_h, num_tiles;
    int act_e, act_dw, has_res, ksteps1;
    int kb, a1_kb_bytes, w1_bytes;  // K blocks of MMA1, bytes of one staged A1 K block / of one W1 chunk
    int off_e, off_a2, off_w, w_buf_bytes, off_f32;
    int off_aux, aux_bytes, a2_bytes;
    int step[3];       // gridDim.x split into (tile column, tile row, image) steps
    const float* aux;  // [n_chunks][K*K + 2][64] fp32: depthwise taps, expand bias, depthwise bias (zero padded)
    const float* b2;
    const bf16* res;
    long long ldres;
    long long* gap;  // depthwise-output mode: [N][Cexp] fixed-point (2^-24) pooling sums, or nullptr
    long long* dbg;  // clock64 stamps of CTA 0 / compute thread 0 (16 per chunk) in -DCAB_MB_DEBUG builds
};

long long* g_mb_dbg = nullptr;

// Phase-boundary clock stamps (tools/dbg_mbconv.py): compiled in only with -DCAB_MB_DEBUG
#ifdef CAB_MB_DEBUG
#define MB_STAMP(cond, slot) do { if (cond) p.dbg[slot] = clock64(); } while (0)
#else
#define MB_STAMP(cond, slot) do { (void)(cond); } while (0)
#endif

__device__ __forceinline__ float lds32f(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_f32(uint32_t a, float v) { asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds128f(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// mbarrier wait for long waits: the suspend-time hint lets the hardware park the warp instead of re-issuing the
// try_wait / branch pair every few cycles (those spins compete with the other CTA's compute warps for issue slots).
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "nanosleep.u32 128;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(tc::smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Tile walk of a persistent CTA (tile += gridDim.x) without divisions: the step is pre-split on the host.
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

// bf16x2 pack of (lo, hi) with the expand activation folded in (ReLU rides on the convert instruction).
template <int ACT> __device__ __forceinline__ uint32_t pack_act(float lo, float hi) {
    uint32_t d;
    if constexpr (ACT == CABINET_ACT_RELU) {
        asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    } else {
        if constexpr (ACT == CABINET_ACT_HSWISH) {
            lo *= __saturatef(fmaf(lo, 1.f / 6.f, 0.5f));  // x * relu6(x + 3) / 6
            hi *= __saturatef(fmaf(hi, 1.f / 6.f, 0.5f));
        }
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    }
    return d;
}

// Per-(warp, lane) pooling partial (two floats): parked in the 16 unused skew bytes behind the 128 channel bytes of the
// staged E rows 0..127 (epilogue 1 never writes them; the E region holds >= 128 rows in this mode) -- no extra shared memory.
__device__ __forceinline__ uint32_t gap_slot(uint32_t sE, int warp, int lane) {
    return sE + (warp * 16 + (lane >> 1)) * E_PITCH + 128 + (lane & 1) * 8;
}

// Depthwise over one row segment: lane = (column group, channel pair).  COLS output columns per lane.
template <int K, int S, int COLS, int IWT, int TW>
__device__ __forceinline__ void dw_seg(const MbParams& p, uint32_t sE, uint32_t sA2, uint32_t s_gap, uint32_t s_aux,
                                       int chunk, int lanes_px, int groups, int row, int seg_col0, bool row_valid,
                                       int ow_base, bool want_gap) {
    constexpr int SPAN = (COLS - 1) * S + K;
    const int lane = threadIdx.x & 31;
    const int grp = lane / lanes_px, cl = lane - grp * lanes_px;
    if (grp >= groups) {
        if (want_gap) {
            const uint32_t a = gap_slot(sE, threadIdx.x >> 5, lane);
            sts32f(a, 0.f);
            sts32f(a + 4, 0.f);
        }
        return;
    }
    const int col0 = seg_col0 + grp * COLS;
    // 3x3: nine taps in registers; 5x5: one filter row at a time from the shared-memory aux block (25 resident taps
    // would need 50 of the 96 registers this kernel may use at two CTAs per SM)
    constexpr bool ROWTAPS = K > 3;
    float2 w[ROWTAPS ? K : K * K];
    if constexpr (!ROWTAPS) {
#pragma unroll
        for (int t = 0; t < K * K; ++t) w[t] = tc::lds64f(s_aux + t * 256 + cl * 8);
    }
    const float2 b = tc::lds64f(s_aux + (K * K + 1) * 256 + cl * 8);
    float2 acc[COLS];
#pragma unroll
    for (int r = 0; r < COLS; ++r) acc[r] = b;
    const uint32_t base = sE + ((row * S) * IWT + col0 * S) * E_PITCH + cl * 4;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        if constexpr (ROWTAPS) {
#pragma unroll
            for (int t = 0; t < K; ++t) w[t] = tc::lds64f(s_aux + (ky * K + t) * 256 + cl * 8);
        }
#pragma unroll
        for (int sx = 0; sx < SPAN; ++sx) {
            uint32_t raw;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(raw) : "r"(base + (ky * IWT + sx) * E_PITCH));
            const float2 xv = make_float2(__uint_as_float(raw << 16), __uint_as_float(raw & 0xffff0000u));
#pragma unroll
            for (int r = 0; r < COLS; ++r) {
                const int kx = sx - r * S;  // compile-time after unrolling
                if (kx >= 0 && kx < K) cab_ffma2(acc[r], xv, w[ROWTAPS ? kx : ky * K + kx]);
            }
        }
    }
    // ReLU rides on the bf16 convert; other activations take the generic path (the SE pooling sums exist only in the
    // activation-free depthwise-output mode)
    const bool relu_pack = p.act_dw == CABINET_ACT_RELU;
    float g0 = 0.f, g1 = 0.f;
    if (want_gap && row_valid) {  // SE pools the BN output, i.e. the values BEFORE the activation (mobilenetv3.py:137-143)
#pragma unroll
        for (int r = 0; r < COLS; ++r) {
            if (ow_base + col0 + r < p.OW) {
                g0 += acc[r].x;
                g1 += acc[r].y;
            }
        }
    }
    if (!relu_pack) cab_act_vec<2 * COLS>(&acc[0].x, p.act_dw);
    const int prow = row * TW + col0;
    const uint32_t a2row = sA2 + prow * 128 + ((cl & 3) << 2);
    const uint32_t cq = static_cast<uint32_t>(cl >> 2) << 4;
#pragma unroll
    for (int r = 0; r < COLS; ++r) {
        // A2 row prow + r, 16-byte chunk (cl >> 2) ^ (row & 7); with 8-aligned segments (prow + r) & 7 == r & 7
        const int sw = (COLS % 8 == 0 && TW % 8 == 0) ? (r & 7) : ((prow + r) & 7);
        const uint32_t hv = relu_pack ? pack_act<CABINET_ACT_RELU>(acc[r].x, acc[r].y)
                                      : pack_act<CABINET_ACT_NONE>(acc[r].x, acc[r].y);
        sts32(a2row + r * 128 + (cq ^ (static_cast<uint32_t>(sw) << 4)), hv);
    }
    if (want_gap) {  // summed over warps / column groups in a fixed order after the A2 barrier
        const uint32_t a = gap_slot(sE, threadIdx.x >> 5, lane);
        sts32f(a, g0);
        sts32f(a + 4, g1);
    }
}

// Epilogue 1 for one TMEM piece (this thread's pixel row, 32 columns): act, bf16, -> E (the bias is already in D1).
template <int ACT>
__device__ __forceinline__ void epi1_piece(const uint32_t* v, uint32_t dst, int ncol, bool keep) {
#pragma unroll
    for (int h16 = 0; h16 < 2; ++h16) {
        if (h16 * 16 < ncol) {
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                o[j] = pack_act<ACT>(__uint_as_float(v[h16 * 16 + 2 * j]), __uint_as_float(v[h16 * 16 + 2 * j + 1]));
            if (!keep) {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = 0u;
            }
            tc::sts128(dst + h16 * 32, make_uint4(o[0], o[1], o[2], o[3]));
            tc::sts128(dst + h16 * 32 + 16, make_uint4(o[4], o[5], o[6], o[7]));
        }
    }
}

__device__ __forceinline__ void bulk_load_1d(uint8_t* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tc::smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}

template <int K, int S, int TH, int TW, bool PROJECT>
__global__ void __launch_bounds__(NTHREADS, 2)
mbconv_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY, const MbParams p) {
    constexpr int PAD = (K - 1) / 2;
    constexpr int IWT = (TW - 1) * S + K, IHT = (TH - 1) * S + K, NPIX = IWT * IHT, NMT = (NPIX + 127) / 128;
    constexpr int NSEG = NCW / TH, WSEG = TW / NSEG, NOUT = TH * TW;
    constexpr int D2COL = NMT * 64;
    static_assert(NCW % TH == 0 && TW % NSEG == 0 && WSEG % 4 == 0 && NOUT <= 128 && D2COL + 64 <= TMEM_COLS, "tile shape");

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a1_full, a1_free, w_full[2], d1_full, d1_free, dw_done, a2_free, d2_full, d2_free;
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sA1 = tc::smem_u32(smem);
    const uint32_t sE = sA1 + p.off_e, sA2 = sA1 + p.off_a2, sW = sA1 + p.off_w;
    const uint32_t s_b2 = sA1 + p.off_f32;                      // [cout_pad] fp32
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nc = p.n_chunks;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmX);
        tc::prefetch_tmap(&tmW1);
        tc::prefetch_tmap(&tmY);
        if (PROJECT) tc::prefetch_tmap(&tmW2);
        tc::mbar_init(&a1_full, 1);
        tc::mbar_init(&a1_free, 1);
        tc::mbar_init(&w_full[0], 1);
        tc::mbar_init(&w_full[1], 1);
        tc::mbar_init(&d1_full, 1);
        tc::mbar_init(&d1_free, NCW);
        tc::mbar_init(&dw_done, 1);
        tc::mbar_init(&a2_free, 1);
        tc::mbar_init(&d2_full, 1);
        tc::mbar_init(&d2_free, NCW);
        tc::mbar_fence_init();
        tc::fence_proxy_async();
    }
    if (warp == NCW) tc::tmem_alloc(&tmem_base_smem, TMEM_COLS);
    if (PROJECT)
        for (int i = threadIdx.x; i < p.cout_pad; i += NTHREADS) sts32f(s_b2 + 4 * i, i < p.Cout ? __ldg(p.b2 + i) : 0.f);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_smem;
    const int n_my = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int exp16 = (p.Cexp + 15) & ~15;

    if (warp == NCW) {
        // ============================ control warp: TMA + MMA issue ============================
        const uint32_t leader = tc::elect_one();
        const uint32_t idesc1 = tc::make_idesc_bf16(128, 64);
        const uint32_t idesc2 = tc::make_idesc_bf16(128, p.cout_pad);
        const uint64_t a1_desc = tc::make_desc_sw128(sA1), a2_desc = tc::make_desc_sw128(sA2);
        const int G = n_my * nc;
        TileIter lt;  // tile whose input patch is loaded next
        lt.init(blockIdx.x, p.tiles_w, p.tiles_h);
        auto load_a1 = [&](int) {
            if (lane == 0) {
                tc::mbar_expect_tx(&a1_full, NPIX * 128 * p.kb);
                for (int kb = 0; kb < p.kb; ++kb)
                    tc::tma_load_4d(smem + kb * p.a1_kb_bytes, &tmX, &a1_full, kb * 64, lt.tw * TW * S - PAD,
                                    lt.th * TH * S - PAD, lt.n);
            }
            lt.next(p.step, p.tiles_w, p.tiles_h);
            __syncwarp();
        };
        auto load_w = [&](int c, int buf) {
            if (lane == 0) {
                uint8_t* dst = smem + p.off_w + buf * p.w_buf_bytes;
                tc::mbar_expect_tx(&w_full[buf], p.w1_bytes + (PROJECT ? p.cout_pad * 128 : 0) + p.aux_bytes);
                for (int kb = 0; kb < p.kb; ++kb) tc::tma_load_2d(dst + kb * W1_KB, &tmW1, &w_full[buf], kb * 64, c * 64);
                if (PROJECT) tc::tma_load_2d(dst + p.w1_bytes, &tmW2, &w_full[buf], c * 64, 0);
                bulk_load_1d(dst + p.off_aux, reinterpret_cast<const uint8_t*>(p.aux) + static_cast<size_t>(c) * p.aux_bytes,
                             p.aux_bytes, &w_full[buf]);
            }
            __syncwarp();
        };
        auto mma1 = [&](int it, int c) {
            const int g = it * nc + c;
            const int buf = p.resident ? c : (g & 1);
            if (c == 0) {
                tc::mbar_wait(&a1_full, it & 1);
                // bias slots: K columns Cin, Cin+1 of every staged pixel <- 1.0 (one 4-byte store per row, in the row's
                // swizzled 16-byte chunk); this warp also issues the MMAs, so a proxy fence is all the ordering needed
                const uint32_t ones = 0x3F803F80u;
                const uint32_t s_ones = sA1 + (p.Cin >> 6) * p.a1_kb_bytes;
                const int cq = (p.Cin & 63) >> 3;
                for (int r = lane; r < NPIX; r += 32) sts32(s_ones + r * 128 + ((cq ^ (r & 7)) << 4), ones);
                tc::fence_proxy_async();
                __syncwarp();
            }
            tc::mbar_wait(&w_full[buf], p.resident ? 0 : ((g >> 1) & 1));
            mbar_wait_parked(&d1_free, (g & 1) ^ 1);
            tc::tc_fence_after();
            const uint64_t b_desc = tc::make_desc_sw128(sW + buf * p.w_buf_bytes);
#pragma unroll
            for (int m = 0; m < NMT; ++m) {
                const uint64_t ad = a1_desc + static_cast<uint64_t>(m * (16384 >> 4));
                tc::umma_bf16_if(leader, tmem + m * 64, ad, b_desc, idesc1, 0u);
                if (p.ksteps1 > 1) tc::umma_bf16_if(leader, tmem + m * 64, ad + 2, b_desc + 2, idesc1, 1u);
                if (p.ksteps1 > 2) tc::umma_bf16_if(leader, tmem + m * 64, ad + 4, b_desc + 4, idesc1, 1u);
                if (p.ksteps1 > 3) tc::umma_bf16_if(leader, tmem + m * 64, ad + 6, b_desc + 6, idesc1, 1u);
                for (int ks = 4; ks < p.ksteps1; ++ks) {  // further K blocks (Cin > 56)
                    const uint64_t ko = static_cast<uint64_t>((ks & 3) * 2);
                    tc::umma_bf16_if(leader, tmem + m * 64, ad + static_cast<uint64_t>(((ks >> 2) * p.a1_kb_bytes) >> 4) + ko,
                                     b_desc + static_cast<uint64_t>(((ks >> 2) * W1_KB) >> 4) + ko, idesc1, 1u);
                }
            }
            tc::umma_commit_if(leader, &d1_full);
            if (c == nc - 1) tc::umma_commit_if(leader, &a1_free);
        };

        load_a1(0);
        if (p.resident) {
            for (int c = 0; c < nc; ++c) load_w(c, c);
        } else {
            load_w(0, 0);
            load_w(1, 1);
        }
        mma1(0, 0);
        for (int it = 0; it < n_my; ++it) {
            for (int c = 0; c < nc; ++c) {
                const int g = it * nc + c;
                if (c + 1 < nc) mma1(it, c + 1);
                if (c == (nc > 1 ? nc - 2 : 0) && it + 1 < n_my) {
                    // the last MMA1 of this tile has been issued: once it has read A1, fetch the next tile's input
                    tc::mbar_wait(&a1_free, it & 1);
                    load_a1(it + 1);
                }
                mbar_wait_parked(&dw_done, g & 1);
                if (PROJECT) {
                    if (c == 0) tc::mbar_wait(&d2_free, (it & 1) ^ 1);
                    tc::tc_fence_after();
                    const int buf = p.resident ? c : (g & 1);
                    const uint64_t b_desc = tc::make_desc_sw128(sW + buf * p.w_buf_bytes + p.w1_bytes);
                    const int ks2 = min(64, exp16 - c * 64) >> 4;
                    tc::umma_bf16_if(leader, tmem + D2COL, a2_desc, b_desc, idesc2, c > 0 ? 1u : 0u);
                    if (ks2 > 1) tc::umma_bf16_if(leader, tmem + D2COL, a2_desc + 2, b_desc + 2, idesc2, 1u);
                    if (ks2 > 2) tc::umma_bf16_if(leader, tmem + D2COL, a2_desc + 4, b_desc + 4, idesc2, 1u);
                    if (ks2 > 3) tc::umma_bf16_if(leader, tmem + D2COL, a2_desc + 6, b_desc + 6, idesc2, 1u);
                    tc::umma_commit_if(leader, &a2_free);
                    if (c == nc - 1) tc::umma_commit_if(leader, &d2_full);
                }
                if (!p.resident && g + 2 < G) {
                    // the weights / taps of chunk g+2 replace those of chunk g once nothing reads them any more:
                    // MMA2(g) complete (project mode) or the depthwise pass of chunk g finished (dw_done above)
                    if (PROJECT) tc::mbar_wait(&a2_free, g & 1);
                    load_w((c + 2) % nc, g & 1);
                }
            }
            if (it + 1 < n_my) mma1(it + 1, 0);
        }
        __syncwarp();
    } else {
        // ============================ compute warps ============================
        const int q = warp & 3, hcol = warp >> 2;     // epilogue 1/2: TMEM lane quarter, column half
        const int row_w = warp / NSEG, seg = warp % NSEG;
        const bool want_gap = !PROJECT && p.gap != nullptr;
        // this thread's pixel (row, column) inside the staged input patch for every M tile: tile independent
        int pty[NMT], ptx[NMT];
#pragma unroll
        for (int m = 0; m < NMT; ++m) {
            const int r = m * 128 + q * 32 + lane;
            pty[m] = r / IWT;
            ptx[m] = r - pty[m] * IWT;
        }
        // epilogue 2: this thread's output pixel inside the tile
        const int pr = q * 32 + lane;
        const int prow = pr / TW, pcol = pr - prow * TW;
        TileIter ti;
        ti.init(blockIdx.x, p.tiles_w, p.tiles_h);
        int g = 0;
        for (int it = 0; it < n_my; ++it, ti.next(p.step, p.tiles_w, p.tiles_h)) {
            const int ow0 = ti.tw * TW, oh0 = ti.th * TH, n = ti.n;
            const int ih0 = oh0 * S - PAD, iw0 = ow0 * S - PAD;
            const bool border = ih0 < 0 || iw0 < 0 || ih0 + IHT > p.H || iw0 + IWT > p.W;
            for (int c = 0; c < nc; ++c, ++g) {
                const int buf = p.resident ? c : (g & 1);
                const uint32_t s_aux = sW + buf * p.w_buf_bytes + p.off_aux;
                const int v16 = min(64, exp16 - c * 64);
                // ---- epilogue 1: D1 -> E (TMEM loads run one piece ahead of the conversion)
                const bool rec = p.dbg && blockIdx.x == 0 && threadIdx.x == 0 && g < 64;
                MB_STAMP(rec, 16 * g + 0);
                tc::mbar_wait(&w_full[buf], p.resident ? 0 : ((g >> 1) & 1));  // aux block (bias, taps) visible
                tc::mbar_wait(&d1_full, g & 1);
                MB_STAMP(rec, 16 * g + 1);
                tc::tc_fence_after();
                const int ncol = min(32, v16 - hcol * 32);
                if (ncol > 0) {
                    uint32_t va[32], vb[32];
                    const uint32_t t0 = tmem + hcol * 32 + (static_cast<uint32_t>(q * 32) << 16);
                    tc::tmem_ld32(t0, va);
#pragma unroll
                    for (int m = 0; m < NMT; ++m) {
                        if (m * 128 + q * 32 >= NPIX) break;
                        uint32_t* cur = (m & 1) ? vb : va;
                        uint32_t* nxt = (m & 1) ? va : vb;
                        tc::tmem_ld_wait();
                        MB_STAMP(rec && m < 3, 16 * g + 9 + 2 * m);
                        if (m + 1 < NMT && (m + 1) * 128 + q * 32 < NPIX) tc::tmem_ld32(t0 + (m + 1) * 64, nxt);
                        const int r = m * 128 + q * 32 + lane;
                        if (r < NPIX) {
                            bool keep = true;
                            if (border) {
                                const int ih = ih0 + pty[m], iw = iw0 + ptx[m];
                                keep = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
                            }
                            const uint32_t dst = sE + r * E_PITCH + hcol * 64;
                            if (p.act_e == CABINET_ACT_RELU) epi1_piece<CABINET_ACT_RELU>(cur, dst, ncol, keep);
                            else if (p.act_e == CABINET_ACT_HSWISH) epi1_piece<CABINET_ACT_HSWISH>(cur, dst, ncol, keep);
                            else epi1_piece<CABINET_ACT_NONE>(cur, dst, ncol, keep);
                        }
                        MB_STAMP(rec && m < 2, 16 * g + 10 + 2 * m);
                    }
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&d1_free);
                MB_STAMP(rec, 16 * g + 2);
                if (threadIdx.x == 0) tc::bulk_wait_read<0>();  // the TMA store that last read A2 is done with it
                tc::named_bar_sync(1, NCW * 32);                // E complete
                if (PROJECT && g > 0) tc::mbar_wait(&a2_free, (g - 1) & 1);
                MB_STAMP(rec, 16 * g + 3);
                // ---- depthwise: E -> A2
                {
                    const bool row_valid = oh0 + row_w < p.OH;
                    if (v16 > 32)
                        dw_seg<K, S, WSEG, IWT, TW>(p, sE, sA2, 0u, s_aux, c, v16 >> 1, 1, row_w, seg * WSEG, row_valid, ow0, want_gap);
                    else if (v16 > 16)
                        dw_seg<K, S, WSEG / 2, IWT, TW>(p, sE, sA2, 0u, s_aux, c, 16, 2, row_w, seg * WSEG, row_valid, ow0, want_gap);
                    else
                        dw_seg<K, S, WSEG / 4, IWT, TW>(p, sE, sA2, 0u, s_aux, c, 8, 4, row_w, seg * WSEG, row_valid, ow0, want_gap);
                }
                MB_STAMP(rec, 16 * g + 4);
                tc::fence_proxy_async();
                tc::named_bar_sync(1, NCW * 32);                // A2 complete, E free
                MB_STAMP(rec, 16 * g + 5);
                if (want_gap && threadIdx.x >= (NCW - 2) * 32) {
                    // deterministic SE pooling sums: fixed-order fp32 sum over the 8 warps / column groups of this
                    // (tile, chunk), then ONE 64-bit INTEGER atomic per channel (fixed point 2^-24; integer addition is
                    // associative, so the total is independent of the CTA arrival order).  Warps 6-7 do it: thread 0
                    // (dw_done arrive + TMA store issue) is not delayed.
                    const int t = threadIdx.x - (NCW - 2) * 32;
                    const int ch = c * 64 + t;
                    if (ch < p.Cexp) {
                        const int lanes_px = v16 > 32 ? (v16 >> 1) : v16 > 16 ? 16 : 8, ng = v16 > 32 ? 1 : v16 > 16 ? 2 : 4;
                        const int cl = t >> 1, hi = t & 1;
                        float sum = 0.f;
                        if (cl < lanes_px) {
                            for (int w8 = 0; w8 < NCW; ++w8)
                                for (int gq = 0; gq < ng; ++gq) sum += lds32f(gap_slot(sE, w8, gq * lanes_px + cl) + hi * 4);
                        }
                        atomicAdd(reinterpret_cast<unsigned long long*>(p.gap + static_cast<long long>(n) * p.Cexp + ch),
                                  static_cast<unsigned long long>(__float2ll_rn(sum * CABINET_GAP_FIXED_ONE)));
                    }
                }
                if (threadIdx.x == 0) {
                    tc::mbar_arrive(&dw_done);
                    if (!PROJECT) {
                        tc::tma_store_4d(&tmY, smem + p.off_a2, c * 64, ow0, oh0, n);
                        tc::bulk_commit();
                    }
                }
            }
            if (PROJECT) {
                // ---- epilogue 2: D2 -> +bias (+identity) -> bf16 -> staging -> TMA store
                const bool rec2 = p.dbg && blockIdx.x == 0 && threadIdx.x == 0 && g <= 64;
                if (q * 32 < NOUT && hcol * 16 < p.cout_pad) {  // this warp owns rows / columns of D2
                    const int oh = oh0 + prow, ow = ow0 + pcol;
                    const bool valid = oh < p.OH && ow < p.OW;
                    const long long pix = (static_cast<long long>(n) * p.OH + oh) * p.OW + ow;
                    // the identity operand of this thread's first 16 output channels travels while MMA2 finishes
                    uint4 res0 = make_uint4(0, 0, 0, 0), res1 = make_uint4(0, 0, 0, 0);
                    if (p.has_res && valid && hcol * 16 < p.Cout) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.ldres + hcol * 16);
                        res0 = __ldg(rp);
                        if (hcol * 16 + 16 <= p.Cout) res1 = __ldg(rp + 1);
                    }
                    tc::mbar_wait(&d2_full, it & 1);
                    MB_STAMP(rec2, 16 * (g - 1) + 6);
                    tc::tc_fence_after();
                    for (int j16 = hcol; j16 * 16 < p.cout_pad; j16 += 2) {
                        uint32_t v[16];
                        tc::tmem_ld16(tmem + D2COL + j16 * 16 + (static_cast<uint32_t>(q * 32) << 16), v);
                        tc::tmem_ld_wait();
                        float f[16];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 bb = lds128f(s_b2 + (j16 * 16 + 4 * j) * 4);
                            f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bb.x;
                            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bb.y;
                            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bb.z;
                            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bb.w;
                        }
                        const int co0 = j16 * 16;
                        if (p.has_res && valid && co0 < p.Cout) {
                            Vec16<bf16> r0, r1;
                            r0.raw = res0;
                            r1.raw = res1;
                            if (j16 != hcol) {
                                const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.ldres + co0);
                                r0.raw = __ldg(rp);
                                r1.raw = co0 + 16 <= p.Cout ? __ldg(rp + 1) : make_uint4(0, 0, 0, 0);
                            }
                            float rf[16];
                            r0.unpack(rf);
                            r1.unpack(rf + 8);
#pragma unroll
                            for (int j = 0; j < 16; ++j) f[j] += rf[j];
                        }
                        Vec16<bf16> o0, o1;
                        o0.pack(f);
                        o1.pack(f + 8);
                        const int grp = j16 >> 2, cc = (j16 & 3) * 2;
                        const uint32_t rowa = (grp == 0 ? sA2 : sE + (grp - 1) * A2_BYTES) + pr * 128;
                        tc::sts128(rowa + ((cc ^ (pr & 7)) << 4), o0.raw);
                        tc::sts128(rowa + (((cc + 1) ^ (pr & 7)) << 4), o1.raw);
                    }
                    tc::tc_fence_before();
                }
                // warps without a share of D2 arrive at once: they touch neither TMEM nor the staging tile
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&d2_free);
                MB_STAMP(rec2, 16 * (g - 1) + 7);
                tc::fence_proxy_async();
                tc::named_bar_sync(1, NCW * 32);
                MB_STAMP(rec2, 16 * (g - 1) + 8);
                if (threadIdx.x == 0) {
                    for (int grp = 0; grp * 64 < p.cout_pad; ++grp)
                        tc::tma_store_4d(&tmY, smem + (grp == 0 ? p.off_a2 : p.off_e + (grp - 1) * A2_BYTES), grp * 64, ow0, oh0, n);
                    tc::bulk_commit();
                    MB_STAMP(rec2, 16 * (g - 1) + 14);
                }
                if (p.cout_pad > 64) {  // part of the staging lives in E, which the next epilogue 1 overwrites
                    if (threadIdx.x == 0) tc::bulk_wait_read<0>();
                    tc::named_bar_sync(1, NCW * 32);
                }
            }
        }
        if (threadIdx.x == 0) tc::bulk_wait_read<0>();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == NCW) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem, TMEM_COLS);
    }
}

template <int K, int S, int TH, int TW, bool PROJECT>
int launch_mb(const void* x, long long ldx, int N, const void* w1, const void* w2, void* y, long long ldy, int cy,
              MbParams p, cudaStream_t st, bool probe) {
    constexpr int IWT = (TW - 1) * S + K, IHT = (TH - 1) * S + K, NPIX = IWT * IHT, NMT = (NPIX + 127) / 128;
    constexpr int NOUT = TH * TW;
    if (NMT * 64 + (PROJECT ? p.cout_pad : 0) > TMEM_COLS) {
        cabinet_set_error("mbconv_fused: TMEM budget (cout %d)", p.Cout);
        return CABINET_ERR_INVALID;
    }
    p.tiles_w = (p.OW + TW - 1) / TW;
    p.tiles_h = (p.OH + TH - 1) / TH;
    const long long tiles = static_cast<long long>(N) * p.tiles_w * p.tiles_h;
    CAB_REQUIRE(tiles < (1LL << 31), "mbconv_fused: too many tiles");
    p.num_tiles = static_cast<int>(tiles);
    p.a1_kb_bytes = ((NPIX * 128 + 1023) / 1024) * 1024;
    p.w1_bytes = p.kb * W1_KB;
    const int a1_bytes = p.kb * p.a1_kb_bytes;
    int e_bytes = ((NPIX * E_PITCH + 1023) / 1024) * 1024;
    if (PROJECT && p.cout_pad > 64) e_bytes = std::max(e_bytes, ((p.cout_pad + 63) / 64 - 1) * A2_BYTES);
    if (!PROJECT) e_bytes = std::max(e_bytes, ((128 * E_PITCH + 1023) / 1024) * 1024);  // gap_slot parks partials behind rows 0..127
    p.a2_bytes = ((NOUT * 128 + 1023) / 1024) * 1024;  // MMA2 reads 128 rows: the tail runs into the weight buffers
    p.off_e = a1_bytes;
    p.off_a2 = p.off_e + e_bytes;
    p.off_w = p.off_a2 + p.a2_bytes;
    p.aux_bytes = (K * K + 2) * 256;
    p.off_aux = p.w1_bytes + (PROJECT ? p.cout_pad * 128 : 0);
    p.w_buf_bytes = ((p.off_aux + p.aux_bytes + 1023) / 1024) * 1024;
    p.resident = p.n_chunks <= 2 ? 1 : 0;
    p.off_f32 = p.off_w + std::min(p.n_chunks, 2) * p.w_buf_bytes;
    const size_t smem = static_cast<size_t>(std::max(p.off_f32, p.off_a2 + A2_BYTES)) + p.cout_pad * 4 + 1024;
    if (smem > 113 * 1024) {
        cabinet_set_error("mbconv_fused: shared-memory budget (%zu bytes)", smem);
        return CABINET_ERR_INVALID;
    }
    if (probe) return CABINET_OK;

    CUtensorMap tmX, tmW1, tmW2, tmY;
    {
        const uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)N};
        const uint64_t strides[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * p.W, (uint64_t)ldx * 2 * p.W * p.H};
        const uint32_t box[4] = {64, (uint32_t)IWT, (uint32_t)IHT, 1};
        int rc = cab_make_tmap_bf16(&tmX, x, 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)p.kb * 64, (uint64_t)((p.Cexp + 15) / 16 * 16)};
        const uint64_t strides[1] = {(uint64_t)p.kb * 128};
        const uint32_t box[2] = {64, 64};
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
    {
        const uint64_t dims[4] = {(uint64_t)cy, (uint64_t)p.OW, (uint64_t)p.OH, (uint64_t)N};
        const uint64_t strides[3] = {(uint64_t)ldy * 2, (uint64_t)ldy * 2 * p.OW, (uint64_t)ldy * 2 * p.OW * p.OH};
        const uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
        int rc = cab_make_tmap_bf16(&tmY, y, 4, dims, strides, box);
        if (rc) return rc;
    }
    int dev = 0, sms = 148;
    CAB_CUDA(cudaGetDevice(&dev));
    CAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = static_cast<int>(std::min<long long>(tiles, 2LL * sms));
    p.step[0] = grid % p.tiles_w;
    p.step[1] = (grid / p.tiles_w) % p.tiles_h;
    p.step[2] = grid / (p.tiles_w * p.tiles_h);
    static bool attr_done = false;
    if (!attr_done) {
        CAB_CUDA(cudaFuncSetAttribute(mbconv_fused_kernel<K, S, TH, TW, PROJECT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      113 * 1024));
        attr_done = true;
    }
    mbconv_fused_kernel<K, S, TH, TW, PROJECT><<<grid, NTHREADS, smem, st>>>(tmX, tmW1, tmW2, tmY, p);
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}

}  // namespace

// not part of the C-ABI header: profiling hook for tools/dbg_mbconv.py (effective in -DCAB_MB_DEBUG builds only)
extern "C" int cabinet_mbconv_debug(long long* device_stamps) {
    g_mb_dbg = device_stamps;
    return CABINET_OK;
}

extern "C" int cabinet_mbconv_fused(const void* x, long long ldx, int N, int H, int W, int Cin, const void* w_expand,
                                    const float* aux_packed, int Cexp, int act_expand, int k, int stride, int act_dw,
                                    const void* w_project, const float* b_project, int Cout, int residual, void* y,
                                    long long ldy, int OH, int OW, long long* gap_sum, cabinet_stream_t stream) {
    CAB_REQUIRE(x && w_expand && aux_packed && y, "mbconv_fused: null pointer");
    CAB_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2), "mbconv_fused: k must be 3|5 and stride 1|2");
    // the expand bias rides in K slots Cin, Cin + 1 of the last K block: that block must have two free slots
    CAB_REQUIRE(N >= 0 && H > 0 && W > 0 && Cin > 0 && Cin <= 248 && Cin % 8 == 0 && Cin % 64 <= 56 && Cexp > 0 &&
                    Cexp % 8 == 0 && Cexp <= 1024,
                "mbconv_fused: needs Cin %% 8 == 0, Cin %% 64 <= 56, Cin <= 248 and Cexp %% 8 == 0 (got Cin %d, Cexp %d)", Cin, Cexp);
    const int pad = (k - 1) / 2;
    CAB_REQUIRE(OH == (H + 2 * pad - k) / stride + 1 && OW == (W + 2 * pad - k) / stride + 1,
                "mbconv_fused: inconsistent output size");
    CAB_REQUIRE(ldx % 8 == 0 && ldx >= Cin && ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_expand) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(aux_packed) & 15) == 0,
                "mbconv_fused: alignment");
    const bool project = w_project != nullptr;
    if (project) {
        CAB_REQUIRE(b_project && Cout > 0 && Cout <= 160 && ldy >= Cout && (reinterpret_cast<uintptr_t>(w_project) & 15) == 0,
                    "mbconv_fused: project needs a bias and Cout <= 160");
        CAB_REQUIRE(!residual || (stride == 1 && Cin == Cout && Cin % 8 == 0), "mbconv_fused: identity needs stride 1, Cin == Cout");
        CAB_REQUIRE(!gap_sum, "mbconv_fused: pooling sums exist in the depthwise-output mode only");
    } else {
        CAB_REQUIRE(ldy >= Cexp && !residual, "mbconv_fused: depthwise-output mode writes Cexp channels, no identity");
    }
    if (N == 0) return CABINET_OK;
    MbParams p;
    p.H = H; p.W = W; p.OH = OH; p.OW = OW; p.Cin = Cin; p.Cexp = Cexp; p.Cout = project ? Cout : 0;
    p.cout_pad = project ? (Cout + 15) / 16 * 16 : 0;
    p.n_chunks = (Cexp + 63) / 64;
    p.act_e = act_expand; p.act_dw = act_dw; p.has_res = residual ? 1 : 0;
    p.ksteps1 = (Cin + 2 + 15) / 16;  // + the two bias slots
    p.kb = Cin / 64 + 1;
    p.aux = aux_packed; p.b2 = b_project;
    p.res = reinterpret_cast<const bf16*>(x); p.ldres = ldx; p.gap = gap_sum; p.dbg = g_mb_dbg;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cy = project ? Cout : Cexp;
#define CAB_MB(K_, S_, TH_, TW_, PROBE_)                                                                           \
    (project ? launch_mb<K_, S_, TH_, TW_, true>(x, ldx, N, w_expand, w_project, y, ldy, cy, p, st, PROBE_)   \
             : launch_mb<K_, S_, TH_, TW_, false>(x, ldx, N, w_expand, w_project, y, ldy, cy, p, st, PROBE_))
    if (k == 3 && stride == 1) {
        if (CAB_MB(3, 1, 8, 16, true) == CABINET_OK) return CAB_MB(3, 1, 8, 16, false);
        return CAB_MB(3, 1, 8, 8, false);  // 10 x 10 input patch = one M tile: what fits beside several K blocks of A1 / W1
    }
    if (k == 5 && stride == 1) {
        if (CAB_MB(5, 1, 8, 16, true) == CABINET_OK) return CAB_MB(5, 1, 8, 16, false);
        return CAB_MB(5, 1, 8, 8, false);  // narrow tile: 12 x 12 input patch, fits next to streamed project weights
    }
    if (k == 3 && stride == 2) {
        if (CAB_MB(3, 2, 4, 16, true) == CABINET_OK) return CAB_MB(3, 2, 4, 16, false);  // wider tile when it fits
        return CAB_MB(3, 2, 4, 8, false);
    }
    return CAB_MB(5, 2, 4, 8, false);
#undef CAB_MB
}
