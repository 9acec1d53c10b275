// Squeeze-excite pooling sums WITHOUT running the depthwise conv (reference: src/models/mobilenetv3.py:68-83,126-143).
//
// In an SE block the gate needs the global mean of d = dw_kxk(h) + b (h = act(W_e x + b_e), the expanded activation)
// before anything downstream of d can be computed.  The mean of a stride-1 "same" depthwise conv is LINEAR in h:
//
//     sum_{oy,ox} d[c] = HW b[c] + sum_{ky,kx} w[c][ky][kx] * R[c](ky - p, kx - p),
//     R(dy, dx) = sum of h over the input window the tap sees = T - (excluded border rows) - (excluded border columns)
//                 + (their corner overlap),
//
// so all it takes per (image, channel) is the total T of h, the sums of its first / last p rows and columns and the
// 2p x 2p corner values (p = (k-1)/2): 9 numbers for 3x3, 25 for 5x5.  This kernel produces them with the expand GEMM on
// the tensor cores, channel-major like mbconv_t.cu (TMEM lane = expanded channel, TMEM column = pixel): a thread adds
// up the columns of its own lane -- pure register adds, no cross-thread reduction, no atomics.
// Each CTA contracts what it summed with the depthwise taps of its channels and adds the result to the block's pooling
// accumulator gap_sum[n][c] as a 64-bit fixed-point integer -- the same deterministic accumulator the depthwise kernels
// fill, so the gate layers (cabinet_gate_fc) do not care where it came from.  The block itself then runs as ONE kernel
// (cabinet_mbconv_t with se_scale): the expanded tensor of an SE block never reaches HBM either.
//
// Work items are <= 256-pixel boxes: "flat" boxes of whole image rows (-> T), single border rows, single border
// columns.  A unit = (image, part): parts 0..nsplit-1 add up a contiguous range of flat boxes each (contribution: (sum
// of the taps) * partial T), the last part does the 4p border rows / columns (contribution: HW * bias - the taps'
// border corrections).  grid = (CTAs per chunk, chunks of 128 channels): a CTA keeps its chunk's expand weights in
// shared memory and walks units u = blockIdx.x, + gridDim.x, ... (persistent: one set-up per CTA).
#include "tc_common.cuh"

namespace {

constexpr int ES_CW = 16;                         // compute warps: 4 TMEM lane quarters x 4 column quarters
constexpr int ES_TMA = ES_CW, ES_MMA = ES_CW + 1;
constexpr int ES_THREADS = (ES_CW + 2) * 32;
constexpr int ES_W1_KB = 128 * 128;
constexpr int ES_MAX_RING = 4;

struct EsParams {
    int H, W, Cin, Cexp, kb, ksteps1, act_e, pad;
    int nblk, rpb, nsplit, parts, units;
    int a1_kb_bytes, a1_stage_bytes, ring, off_w1;
    const float* aux;   // [chunks][k*k + 1][128] fp32: depthwise taps + bias in TMEM-lane order (cabinet_mbconv_t's aux_t)
    long long* gap;     // [N][Cexp] fixed-point (2^-24) pooling sums of the depthwise BN output
};

__device__ __forceinline__ void es_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ void es_tld16(uint32_t a, float* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(a)
        : "memory");
}
__device__ __forceinline__ float es_tld1(uint32_t a) {
    float v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
template <int N> __device__ __forceinline__ void es_act(float* v, int act) {
    if (act == CABINET_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (act == CABINET_ACT_HSWISH) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = v[i] * __saturatef(fmaf(v[i], 1.f / 6.f, 0.5f));
    }
}

// One work item: which tensor map, box origin, pixels in the box, how many of them lie inside the image (row-major
// prefix), MMA N.
struct EsItem {
    int type;       // 0 flat rows, 1 border row, 2 border column
    int x, y;       // box origin
    int npx, nvalid, ncols;
};

__device__ __forceinline__ EsItem es_item(const EsParams& p, int part, int i, int b0) {
    EsItem it;
    const int P = p.pad;
    if (part < p.nsplit) {
        const int y0 = (b0 + i) * p.rpb;
        it.type = 0; it.x = 0; it.y = y0;
        it.npx = p.rpb * p.W;
        it.nvalid = min(p.rpb, p.H - y0) * p.W;
    } else if (i < 2 * P) {
        it.type = 1; it.x = 0; it.y = i < P ? i : p.H - 2 * P + i;
        it.npx = it.nvalid = p.W;
    } else {
        const int j = i - 2 * P;
        it.type = 2; it.x = j < P ? j : p.W - 2 * P + j; it.y = 0;
        it.npx = it.nvalid = p.H;
    }
    it.ncols = (it.npx + 15) & ~15;
    return it;
}

// Position in a CTA's item stream: unit u = (image n, part), item t of it.
struct EsWalk {
    int u, t, n, part, b0, n_items;
    __device__ __forceinline__ void set_unit(const EsParams& p, int u_) {
        u = u_;
        t = 0;
        n = u / p.parts;
        part = u - n * p.parts;
        if (part == p.nsplit) {
            b0 = 0;
            n_items = 4 * p.pad;
        } else {
            b0 = static_cast<int>(static_cast<long long>(part) * p.nblk / p.nsplit);
            n_items = static_cast<int>(static_cast<long long>(part + 1) * p.nblk / p.nsplit) - b0;
        }
    }
    __device__ __forceinline__ bool valid(const EsParams& p) const { return u < p.units; }
    __device__ __forceinline__ void next(const EsParams& p) {
        if (++t >= n_items) set_unit(p, u + static_cast<int>(gridDim.x));
    }
};

template <int P>
__global__ void __launch_bounds__(ES_THREADS, 1)
expand_sums_kernel(const __grid_constant__ CUtensorMap tmFlat, const __grid_constant__ CUtensorMap tmRow,
                   const __grid_constant__ CUtensorMap tmCol, const __grid_constant__ CUtensorMap tmW1, const EsParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a1_full[ES_MAX_RING], a1_ready[ES_MAX_RING], a1_free[ES_MAX_RING], w1_full, d_full[2], d_free[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_part[3][128];  // partial sums of column quarters 1..3, per TMEM lane
    __shared__ float s_border[2 * 2 * P + 4 * P * P][128];  // border row / column sums + corner values, per TMEM lane

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sA1 = tc::smem_u32(smem), sW1 = sA1 + p.off_w1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.y;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmFlat);
        tc::prefetch_tmap(&tmRow);
        tc::prefetch_tmap(&tmCol);
        tc::prefetch_tmap(&tmW1);
        for (int i = 0; i < ES_MAX_RING; ++i) {
            tc::mbar_init(&a1_full[i], 1);
            tc::mbar_init(&a1_ready[i], 1);
            tc::mbar_init(&a1_free[i], 1);
        }
        tc::mbar_init(&w1_full, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&d_full[i], 1);
            tc::mbar_init(&d_free[i], ES_CW);
        }
        tc::mbar_fence_init();
        tc::fence_proxy_async();
    }
    if (warp == ES_MMA) tc::tmem_alloc(&tmem_base_smem, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_smem;

    if (warp == ES_TMA) {
        // ============================ TMA producer + bias slots ============================
        if (lane == 0) {
            tc::mbar_expect_tx(&w1_full, p.kb * ES_W1_KB);
            for (int kb = 0; kb < p.kb; ++kb) tc::tma_load_2d(smem + p.off_w1 + kb * ES_W1_KB, &tmW1, &w1_full, kb * 64, chunk * 128);
        }
        EsWalk wi, wf;  // next item to load / to finalize
        wi.set_unit(p, blockIdx.x);
        wf.set_unit(p, blockIdx.x);
        int issued = 0;
        auto issue = [&]() {
            const int s = issued % p.ring;
            if (issued >= p.ring) tc::mbar_wait(&a1_free[s], ((issued / p.ring) - 1) & 1);
            if (lane == 0) {
                const EsItem it = es_item(p, wi.part, wi.t, wi.b0);
                const CUtensorMap* tm = it.type == 0 ? &tmFlat : (it.type == 1 ? &tmRow : &tmCol);
                uint8_t* dst = smem + s * p.a1_stage_bytes;
                tc::mbar_expect_tx(&a1_full[s], it.npx * 128 * p.kb);
                for (int kb = 0; kb < p.kb; ++kb) tc::tma_load_4d(dst + kb * p.a1_kb_bytes, tm, &a1_full[s], kb * 64, it.x, it.y, wi.n);
            }
            wi.next(p);
            ++issued;
            __syncwarp();
        };
        // ring - 1 loads ahead: the load issued after finalizing item g reuses the stage of item g - 1, whose MMA was kicked
        // off a whole step ago (a full ring would wait for the MMA of the item just finalized: a serial chain)
        const int ahead = p.ring > 1 ? p.ring - 1 : 1;
        while (wi.valid(p) && issued < ahead) issue();
        for (int g = 0; wf.valid(p); ++g, wf.next(p)) {
            const int s = g % p.ring;
            const EsItem it = es_item(p, wf.part, wf.t, wf.b0);
            tc::mbar_wait(&a1_full[s], (g / p.ring) & 1);
            // bias slots (K columns Cin, Cin + 1 <- 1.0) of the pixels inside the image only: the rows of a flat box
            // below the image stay all zero, so they contribute act(0) = 0 to the sums
            const uint32_t s_ones = sA1 + s * p.a1_stage_bytes + (p.Cin >> 6) * p.a1_kb_bytes;
            const int cq = (p.Cin & 63) >> 3;
            for (int r = lane; r < it.nvalid; r += 32) es_sts32(s_ones + r * 128 + ((cq ^ (r & 7)) << 4), 0x3F803F80u);
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&a1_ready[s]);
            if (wi.valid(p)) issue();
        }
        __syncwarp();
    } else if (warp == ES_MMA) {
        // ============================ MMA issue ============================
        const uint32_t leader = tc::elect_one();
        tc::mbar_wait(&w1_full, 0);
        EsWalk wm;
        wm.set_unit(p, blockIdx.x);
        for (int t = 0; wm.valid(p); ++t, wm.next(p)) {
            const int s = t % p.ring, ds = t & 1;
            const EsItem it = es_item(p, wm.part, wm.t, wm.b0);
            tc::mbar_wait(&a1_ready[s], (t / p.ring) & 1);
            if (t >= 2) tc::mbar_wait(&d_free[ds], ((t >> 1) - 1) & 1);
            tc::tc_fence_after();
            const uint32_t idesc = tc::make_idesc_bf16(128, it.ncols);
            const uint64_t a_desc = tc::make_desc_sw128(sW1);
            const uint64_t b_desc = tc::make_desc_sw128(sA1 + s * p.a1_stage_bytes);
            for (int ks = 0; ks < p.ksteps1; ++ks) {
                const uint64_t ko = static_cast<uint64_t>((ks & 3) * 2);
                tc::umma_bf16_if(leader, tmem + ds * 256, a_desc + static_cast<uint64_t>(((ks >> 2) * ES_W1_KB) >> 4) + ko,
                                 b_desc + static_cast<uint64_t>(((ks >> 2) * p.a1_kb_bytes) >> 4) + ko, idesc, ks > 0 ? 1u : 0u);
            }
            tc::umma_commit_if(leader, &d_full[ds]);
            tc::umma_commit_if(leader, &a1_free[s]);
        }
        __syncwarp();
    } else {
        // ============================ compute warps: column sums of this thread's lane ============================
        constexpr int K = 2 * P + 1, P2 = 2 * P;
        const int q = warp & 3, cq4 = warp >> 2;  // TMEM lane quarter, column quarter
        const bool lead = cq4 == 0;               // the warp that owns the channel's result
        const int l = q * 32 + lane;
        const int cg = chunk * 128 + l;
        const bool ch_ok = cg < p.Cexp;
        const uint32_t tlane = tmem + (static_cast<uint32_t>(q * 32) << 16);
        // this warp's quarter of item t's columns (t = position in the CTA's item stream): at most 64 columns, all four
        // TMEM loads in flight at once, then activation + adds
        auto item_sum = [&](const EsItem& it, int t) -> float {
            const int ds = t & 1;
            const int qsz = ((it.npx + 3) / 4 + 15) & ~15;  // <= 64
            const int c_begin = min(cq4 * qsz, it.npx), c_end = min(c_begin + qsz, it.npx);
            tc::mbar_wait(&d_full[ds], (t >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t t0 = tlane + ds * 256 + c_begin;
            const int ncol = c_end - c_begin;
            float buf[4][16];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i * 16 < ncol) es_tld16(t0 + i * 16, buf[i]);
            tc::tmem_ld_wait();
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i * 16 < ncol) {
                    float* w = buf[i];
                    es_act<16>(w, p.act_e);
                    if (i * 16 + 16 > ncol) {  // ragged tail: the columns past the box hold stale data
#pragma unroll
                        for (int j = 0; j < 16; ++j) w[j] = i * 16 + j < ncol ? w[j] : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        a0 += w[j];
                        a1 += w[j + 1];
                        a2 += w[j + 2];
                        a3 += w[j + 3];
                    }
                }
            }
            return (a0 + a1) + (a2 + a3);
        };
        auto release = [&](int t) {
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&d_free[t & 1]);
        };
        // depthwise taps + bias of this thread's channel
        const float* ax = p.aux + static_cast<size_t>(chunk) * (K * K + 1) * 128 + l;
        EsWalk wc;
        int g = 0;
        for (wc.set_unit(p, blockIdx.x); wc.valid(p); wc.set_unit(p, wc.u + static_cast<int>(gridDim.x))) {
            const int n = wc.n, part = wc.part, b0 = wc.b0, n_items = wc.n_items;
            const bool border = part == p.nsplit;
            float tot = 0.f;
            for (int t = 0; t < n_items; ++t, ++g) {
                const EsItem it = es_item(p, part, t, b0);
                const float sum = item_sum(it, g);
                if (!border) {
                    tot += sum;
                } else {
                    // border rows 0..p-1, H-p..H-1 (items 0..2p-1) | border columns 0..p-1, W-p..W-1 (items 2p..4p-1) |
                    // the 2p x 2p corner values: kept per TMEM lane in shared memory (written and read by the lead thread)
                    if (!lead) s_part[cq4 - 1][l] = sum;
                    tc::named_bar_sync(1, ES_CW * 32);
                    if (lead) {
                        s_border[t][l] = ((sum + s_part[0][l]) + s_part[1][l]) + s_part[2][l];
                        if (t < P2) {
                            for (int ci = 0; ci < P2; ++ci) {  // corners of a border row: single TMEM columns
                                const int cc = ci < P ? ci : p.W - P2 + ci;
                                float v1 = es_tld1(tlane + (g & 1) * 256 + cc);
                                tc::tmem_ld_wait();
                                es_act<1>(&v1, p.act_e);
                                s_border[2 * P2 + t * P2 + ci][l] = v1;
                            }
                        }
                    }
                    tc::named_bar_sync(1, ES_CW * 32);
                }
                release(g);
            }
            float contrib = 0.f;
            if (!border) {
                if (!lead) s_part[cq4 - 1][l] = tot;
                tc::named_bar_sync(1, ES_CW * 32);
                if (lead) {
                    float wsum = 0.f;
#pragma unroll
                    for (int i = 0; i < K * K; ++i) wsum += __ldg(ax + i * 128);
                    contrib = wsum * (((tot + s_part[0][l]) + s_part[1][l]) + s_part[2][l]);
                }
                tc::named_bar_sync(1, ES_CW * 32);  // s_part is reused by the next unit
            } else if (lead) {
                // sum over the taps of w * (rows the tap never reads + columns it never reads - their overlap)
                float corr = 0.f;
#pragma unroll
                for (int ky = 0; ky < K; ++ky) {
                    const int dy = ky - P;
                    const int r0 = dy < 0 ? P2 + dy : 0, r1 = dy > 0 ? dy : (dy < 0 ? P2 : 0);
#pragma unroll
                    for (int kx = 0; kx < K; ++kx) {
                        const int dx = kx - P;
                        const int c0 = dx < 0 ? P2 + dx : 0, c1 = dx > 0 ? dx : (dx < 0 ? P2 : 0);
                        float ex = 0.f;
                        for (int r = r0; r < r1; ++r) ex += s_border[r][l];
                        for (int c = c0; c < c1; ++c) ex += s_border[P2 + c][l];
                        for (int r = r0; r < r1; ++r)
                            for (int c = c0; c < c1; ++c) ex -= s_border[2 * P2 + r * P2 + c][l];
                        corr = fmaf(__ldg(ax + (ky * K + kx) * 128), ex, corr);
                    }
                }
                contrib = static_cast<float>(p.H) * static_cast<float>(p.W) * __ldg(ax + K * K * 128) - corr;
            }
            if (lead && ch_ok)
                atomicAdd(reinterpret_cast<unsigned long long*>(p.gap + static_cast<long long>(n) * p.Cexp + cg),
                          static_cast<unsigned long long>(__float2ll_rn(contrib * CABINET_GAP_FIXED_ONE)));
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == ES_MMA) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem, 512);
    }
}

}  // namespace

extern "C" int cabinet_expand_sums(const void* x, long long ldx, int N, int H, int W, int Cin, const void* w_expand_t,
                                   const float* aux_t, int Cexp, int act_expand, int k, int nsplit, long long* gap_sum,
                                   cabinet_stream_t stream) {
    CAB_REQUIRE(x && w_expand_t && aux_t && gap_sum, "expand_sums: null pointer");
    CAB_REQUIRE(k == 3 || k == 5, "expand_sums: k must be 3 or 5");
    const int P = (k - 1) / 2;
    CAB_REQUIRE(N >= 0 && N <= 65535 && H >= 2 * P && W >= 2 * P && H <= 256 && W <= 256,
                "expand_sums: needs 2p <= H, W <= 256 (got %d x %d)", H, W);
    CAB_REQUIRE(Cin > 0 && Cin <= 248 && Cin % 8 == 0 && Cin % 64 <= 56 && Cexp > 64 && Cexp % 8 == 0 && Cexp <= 1024,
                "expand_sums: needs Cin %% 8 == 0, Cin %% 64 <= 56, Cin <= 248, 64 < Cexp <= 1024 (got Cin %d, Cexp %d)", Cin, Cexp);
    CAB_REQUIRE(act_expand >= CABINET_ACT_NONE && act_expand <= CABINET_ACT_HSWISH, "expand_sums: activation none / ReLU / hard-swish");
    CAB_REQUIRE(ldx % 8 == 0 && ldx >= Cin && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(w_expand_t) & 15) == 0,
                "expand_sums: alignment");
    if (N == 0) return CABINET_OK;
    EsParams p;
    p.H = H; p.W = W; p.Cin = Cin; p.Cexp = Cexp; p.act_e = act_expand; p.pad = P;
    p.kb = Cin / 64 + 1;
    p.ksteps1 = (Cin + 2 + 15) / 16;
    p.rpb = std::max(1, 256 / W);
    p.nblk = (H + p.rpb - 1) / p.rpb;
    const int nc = (Cexp + 127) / 128;
    int dev = 0, sms = 148;
    CAB_CUDA(cudaGetDevice(&dev));
    CAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int cpc = std::max(1, sms / nc);  // CTAs per chunk: one wave
    if (nsplit <= 0) nsplit = std::max(1, std::min(p.nblk, (2 * cpc + N - 1) / N - 1));  // ~2 units per CTA
    CAB_REQUIRE(nsplit >= 1 && nsplit <= p.nblk, "expand_sums: nsplit must be in [1, %d] (or <= 0: automatic)", p.nblk);
    p.nsplit = nsplit;
    p.parts = nsplit + 1;
    p.units = N * p.parts;
    p.aux = aux_t;
    p.gap = gap_sum;
    p.a1_kb_bytes = 256 * 128;
    p.a1_stage_bytes = p.kb * p.a1_kb_bytes;
    const int w1_bytes = p.kb * ES_W1_KB;
    p.ring = std::min(ES_MAX_RING, (210 * 1024 - w1_bytes - 2048) / p.a1_stage_bytes);  // 227 KB minus 14 KB static
    CAB_REQUIRE(p.ring >= 1, "expand_sums: shared-memory budget");
    p.off_w1 = p.ring * p.a1_stage_bytes;
    const size_t smem = static_cast<size_t>(p.off_w1) + w1_bytes + 1024;

    CUtensorMap tmFlat, tmRow, tmCol, tmW1;
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * W, (uint64_t)ldx * 2 * W * H};
    {
        const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)p.rpb, 1};
        int rc = cab_make_tmap_bf16(&tmFlat, x, 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint32_t box[4] = {64, (uint32_t)W, 1, 1};
        int rc = cab_make_tmap_bf16(&tmRow, x, 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint32_t box[4] = {64, 1, (uint32_t)H, 1};
        int rc = cab_make_tmap_bf16(&tmCol, x, 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t wd[2] = {(uint64_t)p.kb * 64, (uint64_t)nc * 128};
        const uint64_t ws[1] = {(uint64_t)p.kb * 128};
        const uint32_t box[2] = {64, 128};
        int rc = cab_make_tmap_bf16(&tmW1, w_expand_t, 2, wd, ws, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    static bool attr_done = false;
    if (!attr_done) {
        CAB_CUDA(cudaFuncSetAttribute(expand_sums_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        CAB_CUDA(cudaFuncSetAttribute(expand_sums_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        attr_done = true;
    }
    dim3 grid(static_cast<unsigned>(std::min(cpc, p.units)), static_cast<unsigned>(nc));
    if (P == 1)
        expand_sums_kernel<1><<<grid, ES_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tmFlat, tmRow, tmCol, tmW1, p);
    else
        expand_sums_kernel<2><<<grid, ES_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tmFlat, tmRow, tmCol, tmW1, p);
    CAB_LAUNCH_CHECK();
    return CABINET_OK;
}
