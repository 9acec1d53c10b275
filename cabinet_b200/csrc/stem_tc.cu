// Fused network stems on the tensor cores: both convolutions that read the fp32 NCHW input image
//   sb.conv1      7x7 s2 p3, 3 -> 64, BN, ReLU       (src/models/cabinet.py:111, 19-44)
//   mobile stem   3x3 s2 p1, 3 -> 16, BN, HardSwish  (src/models/mobilenetv3.py:86-91,173)
// are ONE implicit GEMM  D[128 px x 80] = A[128 x 192] * W[80 x 192]^T : the 3x3 filter is embedded in the centre of
// a 7x7 one (same stride, same centre), so the image is read from HBM exactly once (it is 12.6 MB/img fp32, the
// largest single read of the network) and both NHWC bf16 outputs are written once.
//
// K layout: k = (c*7 + ky)*8 + kx with kx in 0..7 (kx = 0 has zero weight: the window starts one column left of
// the filter footprint so that the TMA start coordinate is 16-byte aligned), 168 padded to 192 = 3 K-blocks of 64.
// With that order one 16-byte A chunk (8 bf16) is 8 CONSECUTIVE input floats of one image row, so the im2col
// build is 4 LDS.64 + 4 packs + 1 STS.128 per chunk, written directly in the SWIZZLE_128B K-major UMMA layout.
//
// The folded-BN bias rides in the GEMM: K slots 168/169 of A hold 1.0 and the matching weight columns hold the bias
// split into bf16 hi + lo parts (fp32-accurate to 2^-17), so the epilogue is activation + pack only.
//
// Persistent, warp specialised (448 threads with CONV_WARPS = 8):
//   warp 0      TMA producer: weights once; per tile one 3-D fp32 box {40 cols, 21 rows, 3 ch} (OOB = zero padding)
//   warp 1      TMEM allocator + MMA issuer: 12 x tcgen05.mma (M128 N80 K16) per tile, 2 accumulator stages
//   warps 2-9   converters (2 threads per A row): fp32 window -> bf16 im2col A tile in shared memory (2 stages)
//   warps 10-13 epilogue: tcgen05.ld (80 columns at once), ReLU (cols 0-63) / HardSwish (cols 64-79), bf16 pack,
//               swizzled smem staging, two TMA stores per tile (sb 64 ch, stem 16 ch)
#include "tc_common.cuh"

namespace {

constexpr int TH = 8, TW = 16;             // output patch = 128 pixels
constexpr int WIN_H = 21, WIN_W = 40;      // input window rows / cols (fp32)
constexpr int WIN_BYTES = 3 * WIN_H * WIN_W * 4;      // 10080
constexpr int WIN_STRIDE = 10240;                     // stage pitch (128-byte aligned)
constexpr int WIN_STAGES = 3;
constexpr int KBLOCKS = 3;
constexpr int A_KB_BYTES = 128 * 128;                 // one [128 x 64] bf16 K-block
constexpr int A_STAGE_BYTES = KBLOCKS * A_KB_BYTES;   // 49152
constexpr int A_STAGES = 2;
constexpr int NOUT = 80;                              // 64 + 16
constexpr int W_KB_BYTES = NOUT * 128;                // 10240
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 128;                         // TMEM columns per accumulator stage (80 used)
constexpr int CONV_WARPS = 8;                         // converter warps (CONV_WARPS / 4 threads per A row); 16 measured:
                                                      // 230 -> 285 us (more shared-memory contention, 80-register cap spills)
constexpr int CONV_THREADS = 32 * CONV_WARPS;
constexpr int CONV_PARTS = CONV_THREADS / 128;        // threads sharing one A row
constexpr int EPI_WARP0 = 2 + CONV_WARPS;             // first of the 4 epilogue warps
constexpr int NUM_THREADS = 64 + CONV_THREADS + 128;
constexpr int WINB_PITCH = 96;                        // bf16 copy of the window: row pitch in bytes (40 used + pad: rows 2
                                                      // apart land 16 banks apart, so a warp's two pixel rows never collide)
constexpr int WINB_BYTES = 3 * WIN_H * WINB_PITCH;    // 6048
constexpr int WINB_STRIDE = 6144;
constexpr int CSB_BYTES = 128 * 128;                  // staged [128 px x 64 ch] bf16 (swizzled)
constexpr int CST_BYTES = 128 * 32;                   // staged [128 px x 16 ch] bf16 (dense)
constexpr int SMEM_BYTES =
    KBLOCKS * W_KB_BYTES + A_STAGES * A_STAGE_BYTES + WIN_STAGES * WIN_STRIDE + 2 * (CSB_BYTES + CST_BYTES) +
    2 * WINB_STRIDE + 1024;

struct StemParams {
    int N, OH, OW, tiles_w, tiles_h, num_tiles;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmYsb, const __grid_constant__ CUtensorMap tmYst, const StemParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t w_bar, win_full[WIN_STAGES], win_empty[WIN_STAGES], a_full[A_STAGES],
        a_empty[A_STAGES], acc_full[ACC_STAGES], acc_empty[ACC_STAGES];
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;                                    // 3 x [80 x 64] bf16, swizzled
    uint8_t* sA = sW + KBLOCKS * W_KB_BYTES;               // 2 x 3 x [128 x 64] bf16, swizzled (30720 = 30 x 1024)
    uint8_t* sWin = sA + A_STAGES * A_STAGE_BYTES;         // 3 x [3][21][40] fp32
    uint8_t* sCsb = sWin + WIN_STAGES * WIN_STRIDE;        // 2 x staged sb output (1024-aligned: 30720 = 30 x 1024)
    uint8_t* sCst = sCsb + 2 * CSB_BYTES;                  // 2 x staged stem output
    uint8_t* sWinB = sCst + 2 * CST_BYTES;                 // 2 x bf16 copy of the window

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmX);
        tc::prefetch_tmap(&tmW);
        tc::prefetch_tmap(&tmYsb);
        tc::prefetch_tmap(&tmYst);
        tc::mbar_init(&w_bar, 1);
        for (int s = 0; s < WIN_STAGES; ++s) {
            tc::mbar_init(&win_full[s], 1);
            tc::mbar_init(&win_empty[s], CONV_WARPS);
        }
        for (int s = 0; s < A_STAGES; ++s) {
            tc::mbar_init(&a_full[s], CONV_WARPS);
            tc::mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < ACC_STAGES; ++s) {
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], 4);
        }
        tc::mbar_fence_init();
        tc::fence_proxy_async();
    }
    if (warp == 1) tc::tmem_alloc(&tmem_base_smem, ACC_STAGES * ACC_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_smem;

    const int tiles_per_img = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            tc::mbar_expect_tx(&w_bar, KBLOCKS * W_KB_BYTES);
            for (int kb = 0; kb < KBLOCKS; ++kb) tc::tma_load_2d(sW + kb * W_KB_BYTES, &tmW, &w_bar, kb * 64, 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int s = it % WIN_STAGES;
                const uint32_t ph = (it / WIN_STAGES) & 1;
                const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
                const int oh0 = (r / p.tiles_w) * TH, ow0 = (r % p.tiles_w) * TW;
                tc::mbar_wait(&win_empty[s], ph ^ 1);
                tc::mbar_expect_tx(&win_full[s], WIN_BYTES);
                // the innermost TMA coordinate must be 16-byte aligned: start one column left of the 7x7 footprint
                tc::tma_load_3d(sWin + s * WIN_STRIDE, &tmX, &win_full[s], 2 * ow0 - 4, 2 * oh0 - 3, 3 * img);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer: warp-uniform control flow, one elected lane issues =================
        const uint32_t leader = tc::elect_one();
        const uint32_t idesc = tc::make_idesc_bf16(128, NOUT);
        const uint64_t a_desc0 = tc::make_desc_sw128(tc::smem_u32(sA));
        const uint64_t w_desc0 = tc::make_desc_sw128(tc::smem_u32(sW));
        tc::mbar_wait(&w_bar, 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int as = it % A_STAGES, cs = it % ACC_STAGES;
            const uint32_t aph = (it / A_STAGES) & 1, cph = (it / ACC_STAGES) & 1;
            tc::mbar_wait(&acc_empty[cs], cph ^ 1);
            tc::mbar_wait(&a_full[as], aph);
            tc::tc_fence_after();
            const uint64_t a_desc = a_desc0 + static_cast<uint64_t>(as * (A_STAGE_BYTES >> 4));
            const uint32_t d = tmem + cs * ACC_COLS;
#pragma unroll
            for (int kb = 0; kb < KBLOCKS; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::umma_bf16_if(leader, d, a_desc + (kb * (A_KB_BYTES >> 4) + 2 * k),
                                     w_desc0 + (kb * (W_KB_BYTES >> 4) + 2 * k), idesc, (kb | k) ? 1u : 0u);
            tc::umma_commit_if(leader, &a_empty[as]);
            tc::umma_commit_if(leader, &acc_full[cs]);
        }
        __syncwarp();
    } else if (warp < EPI_WARP0) {
        // ================= converters: fp32 window -> swizzled bf16 im2col (2 threads per A row) =================
        const int ct = threadIdx.x - 64;       // 0..CONV_THREADS-1
        const int r = ct & 127, half = ct >> 7;  // A row = output pixel of the patch; half = which chunks (j = half mod CONV_PARTS)
        const int oy = r / TW, ox = r % TW;
        if (half == 0) {
            // K padding chunks 21..23 of K-block 2 never change: chunk 21 = {1, 1, 0...} (the two bias slots), rest 0
            for (int as = 0; as < A_STAGES; ++as) {
                const uint32_t kb2 = tc::smem_u32(sA) + as * A_STAGE_BYTES + 2 * A_KB_BYTES + r * 128;
                tc::sts128(kb2 + ((5 ^ (r & 7)) << 4), make_uint4(0x3F803F80u, 0, 0, 0));
                tc::sts128(kb2 + ((6 ^ (r & 7)) << 4), make_uint4(0, 0, 0, 0));
                tc::sts128(kb2 + ((7 ^ (r & 7)) << 4), make_uint4(0, 0, 0, 0));
            }
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int ws = it % WIN_STAGES, as = it % A_STAGES;
            const uint32_t wph = (it / WIN_STAGES) & 1, aph = (it / A_STAGES) & 1;
            tc::mbar_wait(&win_full[ws], wph);
            const uint32_t win = tc::smem_u32(sWin) + ws * WIN_STRIDE;
            const uint32_t winb = tc::smem_u32(sWinB) + (it & 1) * WINB_STRIDE;
            // pass 1: every window element is rounded to bf16 ONCE (it is used by ~12 im2col chunks); 1260 float2 pairs
            // over 256 threads.  The im2col pass then moves 16 instead of 32 bytes per chunk and does no conversions:
            // the kernel is shared-memory-bandwidth bound, this removes a third of the converter wavefronts.
#pragma unroll
            for (int k = 0; k < (3 * WIN_H * (WIN_W / 2) + CONV_THREADS - 1) / CONV_THREADS; ++k) {
                const int pidx = ct + CONV_THREADS * k;
                if (pidx < 3 * WIN_H * (WIN_W / 2)) {
                    const float2 f = tc::lds64f(win + pidx * 8);
                    const int rowi = pidx / (WIN_W / 2), cp = pidx - rowi * (WIN_W / 2);
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(winb + rowi * WINB_PITCH + cp * 4), "r"(pack_bf16x2(f.x, f.y)));
                }
            }
            tc::named_bar_sync(2, CONV_THREADS);  // the bf16 window is complete (and the fp32 stage is no longer needed)
            if (lane == 0) tc::mbar_arrive(&win_empty[ws]);
            tc::mbar_wait(&a_empty[as], aph ^ 1);
            const uint32_t arow = tc::smem_u32(sA) + as * A_STAGE_BYTES + r * 128;
#pragma unroll
            for (int jj = 0; jj < (21 + CONV_PARTS - 1) / CONV_PARTS; ++jj) {  // j = c*7 + ky; j = half, half + CONV_PARTS, ...
                const int j = CONV_PARTS * jj + half;
                if (j >= 21) break;
                const int c = j / 7, ky = j % 7;
                const uint32_t src = winb + (c * WIN_H + 2 * oy + ky) * WINB_PITCH + ox * 4;
                uint4 v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.x) : "r"(src));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.y) : "r"(src + 4));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.z) : "r"(src + 8));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.w) : "r"(src + 12));
                const int kb = j >> 3, cc = j & 7;
                tc::sts128(arow + kb * A_KB_BYTES + ((cc ^ (r & 7)) << 4), v);
            }
            tc::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&a_full[as]);
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const bool leader = warp == EPI_WARP0 && lane == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int cs = it % ACC_STAGES;
            const uint32_t cph = (it / ACC_STAGES) & 1;
            const int img = tile / tiles_per_img, rr = tile - img * tiles_per_img;
            const int oh0 = (rr / p.tiles_w) * TH, ow0 = (rr % p.tiles_w) * TW;
            uint8_t* bsb = sCsb + (it & 1) * CSB_BYTES;
            uint8_t* bst = sCst + (it & 1) * CST_BYTES;
            tc::mbar_wait(&acc_full[cs], cph);
            tc::tc_fence_after();
            const uint32_t taddr = tmem + cs * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
            uint32_t a0[32], a1[32], a2[16];
            tc::tmem_ld32(taddr, a0);
            tc::tmem_ld32(taddr + 32, a1);
            tc::tmem_ld16(taddr + 64, a2);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[cs]);   // accumulator stage is free again
            // the stores that used this staging pair two tiles ago must have finished reading it
            if (leader) tc::bulk_wait_read<1>();
            tc::named_bar_sync(1, 128);
            const uint32_t rowp = tc::smem_u32(bsb) + r * 128;
#pragma unroll
            for (int g = 0; g < 8; ++g) {  // 8 channels (16 bytes) per chunk, ReLU
                uint32_t w[4];  // ReLU rides on the convert (cvt.rn.relu.bf16x2)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float lo = __uint_as_float(g < 4 ? a0[8 * g + 2 * j] : a1[8 * (g - 4) + 2 * j]);
                    const float hi = __uint_as_float(g < 4 ? a0[8 * g + 2 * j + 1] : a1[8 * (g - 4) + 2 * j + 1]);
                    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w[j]) : "f"(hi), "f"(lo));
                }
                tc::sts128(rowp + ((g ^ (r & 7)) << 4), make_uint4(w[0], w[1], w[2], w[3]));
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {  // backbone stem: HardSwish
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float t = __uint_as_float(a2[8 * g + j]);
                    v[j] = t * __saturatef(fmaf(t, 1.f / 6.f, 0.5f));
                }
                Vec16<bf16> o;
                o.pack(v);
                tc::sts128(tc::smem_u32(bst) + r * 32 + g * 16, o.raw);
            }
            tc::fence_proxy_async();
            tc::named_bar_sync(1, 128);
            if (leader) {
                tc::tma_store_4d(&tmYsb, bsb, 0, ow0, oh0, img);
                tc::tma_store_4d(&tmYst, bst, 0, ow0, oh0, img);
                tc::bulk_commit();
            }
        }
        if (leader) tc::bulk_wait_read<0>();  // smem must outlive the reads; global visibility comes with grid completion
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem, ACC_STAGES * ACC_COLS);
    }
}

int g_attr_set = 0;

}  // namespace

extern "C" int cabinet_stem_tc(const float* x, int N, int H, int W, const void* w_packed, const float* /*bias: folded*/,
                               void* y_sb, long long ld_sb, void* y_stem, long long ld_stem, int OH, int OW,
                               cabinet_stream_t stream) {
    CAB_REQUIRE(x && w_packed && y_sb && y_stem, "stem_tc: null pointer");
    CAB_REQUIRE(N >= 0 && H > 0 && W > 0 && W % 4 == 0, "stem_tc: W must be a multiple of 4 (TMA row pitch)");
    CAB_REQUIRE(OH == (H - 1) / 2 + 1 && OW == (W - 1) / 2 + 1, "stem_tc: inconsistent output size");
    CAB_REQUIRE(ld_sb >= 64 && ld_sb % 8 == 0 && ld_stem >= 16 && ld_stem % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(y_sb) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_stem) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
                "stem_tc: alignment");
    if (N == 0) return CABINET_OK;
    cab_encode_tiled_fn enc = cab_get_encode_tiled();
    CAB_REQUIRE(enc != nullptr, "stem_tc: cuTensorMapEncodeTiled unavailable");

    CUtensorMap tmX, tmW;
    {
        cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)3 * N};
        cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
        cuuint32_t box[3] = {WIN_W, WIN_H, 3};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            cabinet_set_error("stem_tc: input tensor map encode failed (CUresult %d)", (int)r);
            return CABINET_ERR_CUDA;
        }
    }
    {
        const uint64_t dims[2] = {KBLOCKS * 64, NOUT};
        const uint64_t strides[1] = {KBLOCKS * 64 * 2};
        const uint32_t box[2] = {64, NOUT};
        int rc = cab_make_tmap_bf16(&tmW, w_packed, 2, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    StemParams p;
    p.N = N; p.OH = OH; p.OW = OW;
    p.tiles_w = (OW + TW - 1) / TW;
    p.tiles_h = (OH + TH - 1) / TH;
    const long long nt = static_cast<long long>(N) * p.tiles_w * p.tiles_h;
    CAB_REQUIRE(nt < (1LL << 31), "stem_tc: too many tiles");
    p.num_tiles = static_cast<int>(nt);
    CUtensorMap tmYsb, tmYst;
    {
        const uint64_t dsb[4] = {64, (uint64_t)OW, (uint64_t)OH, (uint64_t)N};
        const uint64_t ssb[3] = {(uint64_t)ld_sb * 2, (uint64_t)ld_sb * 2 * OW, (uint64_t)ld_sb * 2 * OW * OH};
        const uint32_t bsb[4] = {64, TW, TH, 1};
        int rc = cab_make_tmap_bf16(&tmYsb, y_sb, 4, dsb, ssb, bsb);
        if (rc) return rc;
        const uint64_t dst[4] = {16, (uint64_t)OW, (uint64_t)OH, (uint64_t)N};
        const uint64_t sst[3] = {(uint64_t)ld_stem * 2, (uint64_t)ld_stem * 2 * OW, (uint64_t)ld_stem * 2 * OW * OH};
        const uint32_t bst[4] = {16, TW, TH, 1};
        rc = cab_make_tmap_bf16(&tmYst, y_stem, 4, dst, sst, bst, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    if (!g_attr_set) {
        CAB_CUDA(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        g_attr_set = 1;
    }
    int dev = 0, sms = 148;
    CAB_CUDA(cudaGetDevice(&dev));
    CAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = static_cast<int>(std::min<long long>(nt, sms));
    stem_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(tmX, tmW, tmYsb, tmYst, p);
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}
