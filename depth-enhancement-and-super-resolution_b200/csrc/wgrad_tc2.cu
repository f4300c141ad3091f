// Second-generation weight-gradient GEMM (sm_100a, tcgen05 / TMEM / TMA).
//
//     dWp[cm][t*Ca + c] (+)= sum over the pixels p of the base grid of  M[p + moff][cm] * A[p + aoff + tap_t][c]
//
// Same contraction, operands and packed output as wgrad_tc_kernel (conv_tc.cu), different schedule.  Profiling the first
// kernel (profiles/r1d_full_wgrad_tc_kernel.md, r1e_traffic.json) showed it bound by the L2 -> SM operand feed and by fixed
// per-CTA costs, not by the tensor pipe: one CTA = ONE tap x <= 256 columns, so the 128-row M tile (dY) crossed L2 -> SM once
// per tap (28 times for the 7x7 head: 192 B per MMA clock against ~42 B/clk/SM of L2 bandwidth) and layers with short K ran
// ~300 CTAs of ~20 K steps each behind 33-way split-K scalar atomics.  Here
//   * the columns of dWp are cut into 64-wide COLUMN BLOCKS (tap t, channel block cb) and one CTA accumulates NB = 8 of them
//     at once in all 512 TMEM columns: per K tile the M operand is loaded once and multiplied with 8 A boxes (taps and / or
//     channel blocks), i.e. 4x .. 8x fewer M bytes per MMA and 8x fewer CTAs / epilogues / atomics;
//   * four column blocks sit LBO = one box apart in shared memory, so ONE tcgen05.mma with N = 256 covers four taps;
//   * the K tile is 32 pixels (4 KB boxes, 40 KB stages, 4 stages in flight);
//   * split-K partial sums leave as 16-byte vector reductions (red.global.add.v4.f32), a quarter of the atomic operations.
// Both operands are MN-major (the pixel dimension is the MMA K dimension): SWIZZLE_128B boxes of [pixels][64 channels],
// descriptor SBO = 1024 B (next 8 pixels), LBO = one box (next 64 channels or next column block).
// Warp roles (192 threads): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue (TMEM lane quadrant = warp % 4).
#include <stdlib.h>
#include "tc_common.cuh"

#define WG2_MAX_TAPS 64

struct Wg2Params {
    int N, Hb, Wb;            // base grid
    int TW, TH, TN;           // K tile = TW x TH x TN pixels
    int tiles_w, tiles_h, tiles_total, tiles_per_split;
    int mh, mw, ah, aw;       // coordinate offsets of the two operands
    int Cm_real, ncb, nblocks, ld;   // rows of dWp, 64-channel blocks per tap, column blocks in total (= T * ncb), row length
    int per;                  // column blocks per CTA (<= NB; the groups are balanced: per = ceil(nblocks / groups))
    int f16;
    float out_scale;
    int partial;              // 1: every K split STORES its own [Cm_real][ld] slab (dWp + z * split_stride), no atomics
    long split_stride;
    signed char dr[WG2_MAX_TAPS], ds[WG2_MAX_TAPS];
};

template <int NB, int KT>
struct Wg2Cfg {
    static constexpr int BOX = KT * 128;                     // KT pixels x 64 channels x 2 B
    static constexpr int STAGE = (2 + NB) * BOX;             // M: 128 channels = 2 boxes; A: NB column blocks
    static constexpr int STAGES = (192 * 1024 / STAGE) > 6 ? 6 : (192 * 1024 / STAGE);
    static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
    static constexpr int TMEM_COLS = NB * 64;
};

__device__ __forceinline__ uint64_t wg2_sdesc_mn(uint32_t lbo_bytes) {
    return ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int NB, int KT>
__global__ void __launch_bounds__(192, 1)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap mapM, const __grid_constant__ CUtensorMap mapA,
                 const __grid_constant__ Wg2Params p, float* __restrict__ dWp) {
    using Cfg = Wg2Cfg<NB, KT>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * Cfg::STAGES);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * Cfg::STAGES + 1);
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int gb0 = blockIdx.x * p.per;                       // first column block of this CTA
    int nb = p.nblocks - gb0;                                 // column blocks that exist (the last group may be short)
    if (nb > p.per) nb = p.per;
    const int cm0 = blockIdx.y * 128;
    const int kb = blockIdx.z * p.tiles_per_split;
    int ke = kb + p.tiles_per_split;
    if (ke > p.tiles_total) ke = p.tiles_total;
    const int nk = ke - kb;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapM); tma_prefetch_desc(&mapA);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_ptr_gen;

    if (warp == 0) {
        // ===== TMA producer =====
        const uint32_t stage_tx = (uint32_t)(2 + nb) * Cfg::BOX;
        int s = 0;
        uint32_t ph = 1u;
        // K-tile coordinates and the first column block advance incrementally (no per-stage divisions)
        int tw_i = kb % p.tiles_w, th_i = (kb / p.tiles_w) % p.tiles_h, n_i = kb / (p.tiles_w * p.tiles_h);
        const int t0 = gb0 / p.ncb, cb0 = gb0 - t0 * p.ncb;
        if (elect_one())                 // one thread runs the whole role (conv_tc3.cu: no per-stage ELECT / reconvergence)
        for (int i = 0; i < nk; ++i) {
            mbar_wait(empty_bar(s), ph);
            {
                const int n0 = n_i * p.TN, h0 = th_i * p.TH, w0 = tw_i * p.TW;
                uint32_t dst = smem_base + s * Cfg::STAGE;
                mbar_expect_tx(full_bar(s), stage_tx);
                tma_load_4d(dst, &mapM, full_bar(s), cm0, w0 + p.mw, h0 + p.mh, n0);
                tma_load_4d(dst + Cfg::BOX, &mapM, full_bar(s), cm0 + 64, w0 + p.mw, h0 + p.mh, n0);
                dst += 2 * Cfg::BOX;
                int t = t0, cb = cb0;
                for (int b = 0; b < nb; ++b, dst += Cfg::BOX) {
                    tma_load_4d(dst, &mapA, full_bar(s), cb << 6, w0 + p.aw + p.ds[t], h0 + p.ah + p.dr[t], n0);
                    if (++cb == p.ncb) { cb = 0; ++t; }
                }
            }
            if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
            if (++tw_i == p.tiles_w) { tw_i = 0; if (++th_i == p.tiles_h) { th_i = 0; ++n_i; } }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: both operands MN-major (instruction-descriptor bits 15 / 16); up to two MMAs of N <= 256 per K step
        const uint32_t fm = (p.f16 & 1) ? 0u : 1u, fa = (p.f16 & 2) ? 0u : 1u;
        const int n_lo = (nb > 4 ? 4 : nb) * 64, n_hi = (nb > 4 ? nb - 4 : 0) * 64;
        const uint32_t ibase = (1u << 4) | (fm << 7) | (fa << 10) | ((uint32_t)(128 >> 4) << 24) | (1u << 15) | (1u << 16);
        const uint32_t idesc_lo = ibase | ((uint32_t)(n_lo >> 3) << 17), idesc_hi = ibase | ((uint32_t)(n_hi >> 3) << 17);
        const uint64_t desc0 = wg2_sdesc_mn(Cfg::BOX);
        int s = 0;
        uint32_t ph = 0u;
        if (elect_one())
        for (int i = 0; i < nk; ++i) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            {
                const uint32_t st = smem_base + s * Cfg::STAGE;
                const uint64_t md = desc0 + (uint64_t)((st & 0x3FFFF) >> 4);
                const uint64_t ad = md + (uint64_t)((2 * Cfg::BOX) >> 4);
#pragma unroll
                for (int kk = 0; kk < KT / 16; ++kk) {
                    tc_mma_bf16(tmem_acc, md + 128 * kk, ad + 128 * kk, idesc_lo, (i | kk) ? 1u : 0u);
                    if (n_hi) tc_mma_bf16(tmem_acc + 256u, md + 128 * kk, ad + (uint64_t)((4 * Cfg::BOX) >> 4) + 128 * kk, idesc_hi, (i | kk) ? 1u : 0u);
                }
                tc_commit(empty_bar(s));
                if (i == nk - 1) tc_commit(tmem_full_bar);
            }
            if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
        }
    } else {
        // ===== epilogue: TMEM lane = row of dWp, 32 columns per load, 128 contiguous bytes per thread and chunk =====
        const int q = warp & 3;
        const int cm = cm0 + q * 32 + lane;
        const bool valid = cm < p.Cm_real && nk > 0;
        float* orow = dWp + (p.partial ? (long)blockIdx.z * p.split_stride : 0L) + (long)cm * p.ld + (long)gb0 * 64;
        const bool split = gridDim.z > 1 && !p.partial;
        if (nk > 0) {
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < nb * 64; c0 += 32) {
                uint32_t v[32];
                tc_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                tc_wait_ld();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float a = __uint_as_float(v[j]) * p.out_scale, b = __uint_as_float(v[j + 1]) * p.out_scale;
                        const float c = __uint_as_float(v[j + 2]) * p.out_scale, d = __uint_as_float(v[j + 3]) * p.out_scale;
                        if (split) red_add_v4(orow + c0 + j, a, b, c, d);
                        else *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(a, b, c, d);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

template <int NB, int KT>
static int launch_wg2(const CUtensorMap& mm, const CUtensorMap& ma, const Wg2Params& p, float* dWp, dim3 grid, cudaStream_t st) {
    using Cfg = Wg2Cfg<NB, KT>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(wgrad_tc2_kernel<NB, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) {
            dsr_set_error("wgrad_tc2: cannot raise dynamic shared memory to %d", Cfg::SMEM);
            return DSR_ERR_CUDA;
        }
        attr = true;
    }
    wgrad_tc2_kernel<NB, KT><<<grid, 192, Cfg::SMEM, st>>>(mm, ma, p, dWp);
    return dsr_check_launch("wgrad_tc2");
}

// K tiling and split count of a launch (shared by the launch and by dsr_tc_wgrad2_splits)
static void wg2_plan(int N, int Hb, int Wb, int Cm_real, int Ca, int T, int split_k, Wg2Params& p, int& groups, int& tiles_m, int& splits) {
    constexpr int NB = 8, KT = 32;
    p.ncb = Ca / 64; p.nblocks = T * p.ncb; p.ld = T * Ca;
    int TW = Wb >= 16 ? 16 : pow2_ceil(Wb);
    if (TW > KT) TW = KT;
    int TH = KT / TW;
    if (TH > pow2_ceil(Hb)) TH = pow2_ceil(Hb);
    int TN = KT / (TW * TH);
    p.TW = TW; p.TH = TH; p.TN = TN;
    p.tiles_w = dsr_cdiv(Wb, TW); p.tiles_h = dsr_cdiv(Hb, TH);
    p.tiles_total = p.tiles_w * p.tiles_h * dsr_cdiv(N, TN);
    groups = dsr_cdiv(p.nblocks, NB);
    p.per = dsr_cdiv(p.nblocks, groups);
    groups = dsr_cdiv(p.nblocks, p.per);
    tiles_m = dsr_cdiv(Cm_real, 128);
    const long ctas = (long)groups * tiles_m;
    splits = 1;
    if (split_k < 0) {
        // one wave of CTAs over the SMs (1 CTA / SM: 160 KB of stages, all 512 TMEM columns), at least 8 K tiles per split
        splits = (int)(dsr_num_sms() / ctas);
        // at least 32 K tiles per split (DSR_WG2_MIN_TILES for A/B; was 8): every split ends with 128 x per x 64 fp32 vector
        // reductions, and a 128-channel 3x3 layer at 12 x 64^2 cut 48 ways put 29 MB of them on a 590 KB matrix.  Measured
        // (r4o / r4p, same box): step 13.98 ms at 8, 13.92 at 16, 13.88 at 32, 13.8-13.9 at 64, 14.08 at 128.
        static const int min_tiles = getenv("DSR_WG2_MIN_TILES") ? atoi(getenv("DSR_WG2_MIN_TILES")) : 32;
        if (splits > p.tiles_total / min_tiles) splits = p.tiles_total / min_tiles;
        if (splits < 1) splits = 1;
    } else if (split_k > 1) splits = split_k;
    p.tiles_per_split = dsr_cdiv(p.tiles_total, splits);
    splits = dsr_cdiv(p.tiles_total, p.tiles_per_split);
}

// number of K splits dsr_tc_wgrad2 / dsr_tc_wgrad2p will run for this shape (>= 1; < 0 on bad arguments): the caller of the
// partial-slab form sizes its buffer with it
extern "C" int dsr_tc_wgrad2_splits(int N, int Hb, int Wb, int Cm_real, int Ca, int T, int split_k) {
    if (N < 1 || Hb < 1 || Wb < 1 || Cm_real < 1 || T < 1 || T > WG2_MAX_TAPS || (Ca & 63) || Ca < 64) return -1;
    Wg2Params p;
    int groups, tiles_m, splits;
    wg2_plan(N, Hb, Wb, Cm_real, Ca, T, split_k, p, groups, tiles_m, splits);
    return splits;
}

// Single-pass (hi planes only) weight gradient; same arguments as dsr_tc_wgrad minus the low planes.
// partial = 0: dWp is ONE [Cm_real][T*Ca] matrix, K splits add into it with vector reductions (memset first).
// partial = 1: dWp holds dsr_tc_wgrad2_splits(...) slabs of [Cm_real][T*Ca]; split z stores slab z with plain stores and
//              dsr_tc_unpack_wgrad_splits sums the slabs in a fixed order on the way into the parameter layout - no memset, no
//              atomics (a 48-way split of a 3x3 128-channel layer put 28 MB of red.add on a 590 KB matrix), and the weight
//              gradient becomes bit-reproducible.
extern "C" int dsr_tc_wgrad2p(const void* M_hi, int N, int Hm, int Wm, int Cm, int Cm_real, int m_off_h, int m_off_w,
                              const void* A_hi, int Ha, int Wa, int Ca, int T, const int* tap_dr, const int* tap_ds,
                              int a_off_h, int a_off_w, int Hb, int Wb, float* dWp, int f16, float out_scale, int split_k,
                              int partial, void* stream) {
    DSR_REQUIRE(M_hi && A_hi && dWp && tap_dr && tap_ds, "null pointer");
    DSR_REQUIRE(T >= 1 && T <= WG2_MAX_TAPS && (Ca & 63) == 0 && (Cm & 63) == 0 && Cm_real >= 1 && Cm_real <= Cm, "bad GEMM shape");
    DSR_REQUIRE(!((uintptr_t)dWp & 15) && (((long)T * Ca) & 3) == 0, "packed gradient rows must be 16-byte aligned");
    constexpr int NB = 8, KT = 32;
    Wg2Params p;
    p.N = N; p.Hb = Hb; p.Wb = Wb; p.mh = m_off_h; p.mw = m_off_w; p.ah = a_off_h; p.aw = a_off_w;
    p.Cm_real = Cm_real; p.f16 = f16; p.out_scale = out_scale;
    for (int t = 0; t < T; ++t) { p.dr[t] = (signed char)tap_dr[t]; p.ds[t] = (signed char)tap_ds[t]; }
    int groups, tiles_m, splits;
    wg2_plan(N, Hb, Wb, Cm_real, Ca, T, split_k, p, groups, tiles_m, splits);
    p.partial = partial ? 1 : 0;
    p.split_stride = (long)Cm_real * T * Ca;
    const int TW = p.TW, TH = p.TH, TN = p.TN;
    if (!partial && splits > 1 && cudaMemsetAsync(dWp, 0, (size_t)Cm_real * T * Ca * sizeof(float), ST(stream)) != cudaSuccess) {
        dsr_set_error("wgrad_tc2: memset failed"); return DSR_ERR_CUDA;
    }
    CUtensorMap mm, ma;
    cuuint64_t mdims[4] = {(cuuint64_t)Cm, (cuuint64_t)Wm, (cuuint64_t)Hm, (cuuint64_t)N};
    cuuint64_t mstr[3] = {(cuuint64_t)Cm * 2, (cuuint64_t)Wm * Cm * 2, (cuuint64_t)Hm * Wm * Cm * 2};
    cuuint64_t adims[4] = {(cuuint64_t)Ca, (cuuint64_t)Wa, (cuuint64_t)Ha, (cuuint64_t)N};
    cuuint64_t astr[3] = {(cuuint64_t)Ca * 2, (cuuint64_t)Wa * Ca * 2, (cuuint64_t)Ha * Wa * Ca * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
    int rc = encode_map(&mm, M_hi, 4, mdims, mstr, box);
    if (rc) return rc;
    if ((rc = encode_map(&ma, A_hi, 4, adims, astr, box))) return rc;
    dim3 grid((unsigned)groups, (unsigned)tiles_m, (unsigned)splits);
    return launch_wg2<NB, KT>(mm, ma, p, dWp, grid, ST(stream));
}

extern "C" int dsr_tc_wgrad2(const void* M_hi, int N, int Hm, int Wm, int Cm, int Cm_real, int m_off_h, int m_off_w,
                             const void* A_hi, int Ha, int Wa, int Ca, int T, const int* tap_dr, const int* tap_ds,
                             int a_off_h, int a_off_w, int Hb, int Wb, float* dWp, int f16, float out_scale, int split_k,
                             void* stream) {
    return dsr_tc_wgrad2p(M_hi, N, Hm, Wm, Cm, Cm_real, m_off_h, m_off_w, A_hi, Ha, Wa, Ca, T, tap_dr, tap_ds, a_off_h, a_off_w,
                          Hb, Wb, dWp, f16, out_scale, split_k, 0, stream);
}
