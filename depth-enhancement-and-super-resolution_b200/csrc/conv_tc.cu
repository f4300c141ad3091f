// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// Every convolution / transposed convolution of the five networks (models/networks.py:379-415,544-616;
// models/translation_network.py:472-570) is brought to ONE GEMM form by the operand-preparation
// kernel (which also fuses the preceding norm-apply + activation + padding, so the padded /
// re-arranged bf16 operand is written exactly once):
//
//     out[n, h*os+ph, w*os+pw, co] = bias[co] + sum_{t < T} sum_{c < Ca}  A[n, h+ah+dr_t, w+aw+ds_t, c] * W[co][t*Ca + c]
//
//   A  : bf16 NHWC "arranged" activation [N][Ha][Wa][Ca], Ca % 64 == 0, already padded, optionally
//        space-to-depth (stride-2 convs become 2x2 stride-1 taps over 4*C channels) or channel-paired
//        (C = 32: two adjacent pixels side by side so one 128-byte swizzle row holds a full K block);
//   W  : bf16 [Cout][T*Ca], K-major;  transposed convs run as 4 output phases (os = 2).
//   Precision: `npass` = 1 (bf16), 2 (A = hi+lo), 3 (A and W = hi+lo, ~16-bit significands, what the
//   parity gates need - SURVEY.md Appendix E); all passes accumulate into the same fp32 TMEM tile.
//
// Kernel structure (one 128 x BLOCK_N output tile per CTA, optional split-K over gridDim.z):
//   warp 0   : TMA producer  - per K step: 4-D box loads of the A tile(s) (box = 64ch x TW x TH x TN
//              pixels = 128 rows of 128 B, SWIZZLE_128B) and 2-D box loads of the W tile(s)
//   warp 1   : TMEM allocator + MMA issuer - tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16,
//              smem descriptors K-major SWIZZLE_128B; tcgen05.commit releases the smem stage / signals the epilogue
//   warps 2-5: epilogue - tcgen05.ld 32x32b.x32 -> +bias -> (tanh) -> fp32 NHWC stores (or red.add for split-K)
#include <stdlib.h>
#include "tc_common.cuh"

// ------------------------------------------------------------------------------------------------
// GEMM kernel
// ------------------------------------------------------------------------------------------------
#define TC_MAX_TAPS 64
struct TcParams {
    int N, Ht, Wt;            // tile-space output grid (per phase)
    int TW, TH, TN;           // tile = TN x TH x TW pixels = 128 rows
    int tiles_w, tiles_h;     // tile counts along w / h (tiles_n = gridDim.x / (tiles_w * tiles_h))
    int Ca, T;                // arranged channels, taps
    int ah, aw;               // A coordinate offsets
    int Ho, Wo, Cout, os, ph, pw;   // output tensor geometry
    int act, ksteps_per_split, ksteps;
    int nphase, tiles_per_phase;   // nphase = 4: the four output phases of a stride-2 transposed conv in ONE launch
                                   // (phase (a, b): A offsets + (a, b), output offsets (a, b), weight rows + phase*Cout)
    int f16;                  // operands are IEEE half (11-bit significand) instead of bf16
    float out_scale;          // accumulator scale (undoes the power-of-two weight scale of the f16 path)
    signed char dr[TC_MAX_TAPS], ds[TC_MAX_TAPS];
};

template <int BLOCK_N, int NPASS>
struct TcCfg {
    static constexpr int A_TILE = 128 * 128;                 // 128 rows x 128 B
    static constexpr int W_TILE = BLOCK_N * 128;
    static constexpr int NA = NPASS >= 2 ? 2 : 1;
    static constexpr int NW = NPASS >= 3 ? 2 : 1;
    static constexpr int STAGE = NA * A_TILE + NW * W_TILE;
    static constexpr int STAGES = (200 * 1024 / STAGE) > 8 ? 8 : (200 * 1024 / STAGE);
    static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
};

template <int BLOCK_N, int NPASS>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
               const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
               const __grid_constant__ TcParams p, const float* __restrict__ bias, float* __restrict__ out) {
    using Cfg = TcCfg<BLOCK_N, NPASS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE;    // full[S], empty[S], tmem_full, tmem_ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * Cfg::STAGES);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * Cfg::STAGES + 1);
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::STAGES * Cfg::STAGE + 8 * (2 * Cfg::STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    int tile = blockIdx.x;
    const int phase = tile / p.tiles_per_phase;
    tile -= phase * p.tiles_per_phase;
    const int pa = phase >> 1, pb = phase & 1;
    const int tw_i = tile % p.tiles_w; tile /= p.tiles_w;
    const int th_i = tile % p.tiles_h; tile /= p.tiles_h;
    const int n0 = tile * p.TN, h0 = th_i * p.TH, w0 = tw_i * p.TW;
    const int co0 = blockIdx.y * BLOCK_N;
    const int kb = blockIdx.z * p.ksteps_per_split;
    int ke = kb + p.ksteps_per_split;
    if (ke > p.ksteps) ke = p.ksteps;
    const int nk = ke - kb;
    const int cblocks = p.Ca >> 6;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA_hi); tma_prefetch_desc(&mapW_hi);
        if (NPASS >= 2) tma_prefetch_desc(&mapA_lo);
        if (NPASS >= 3) tma_prefetch_desc(&mapW_lo);
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
        // ===== TMA producer (whole warp loops and waits, one elected lane issues) =====
        // one elected thread runs the whole role; ring slot / parity / (tap, channel block) advance incrementally (conv_tc3.cu)
        int s = 0, t = kb / cblocks, cb = kb - t * cblocks;
        uint32_t ph = 1u;
        if (elect_one())
        for (int i = 0; i < nk; ++i) {
            mbar_wait(empty_bar(s), ph);
            {
                const int ca = cb << 6;
                const int hc = h0 + p.ah + pa + p.dr[t], wc = w0 + p.aw + pb + p.ds[t];
                const uint32_t st = smem_base + s * Cfg::STAGE;
                mbar_expect_tx(full_bar(s), Cfg::STAGE);
                tma_load_4d(st, &mapA_hi, full_bar(s), ca, wc, hc, n0);
                if (NPASS >= 2) tma_load_4d(st + Cfg::A_TILE, &mapA_lo, full_bar(s), ca, wc, hc, n0);
                const int kw = t * p.Ca + ca, wrow = phase * p.Cout + co0;
                tma_load_2d(st + Cfg::NA * Cfg::A_TILE, &mapW_hi, full_bar(s), kw, wrow);
                if (NPASS >= 3) tma_load_2d(st + Cfg::NA * Cfg::A_TILE + Cfg::W_TILE, &mapW_lo, full_bar(s), kw, wrow);
            }
            if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
            if (++cb == cblocks) { cb = 0; ++t; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread runs the loop) =====
        const uint32_t idesc = make_idesc(128, BLOCK_N < 16 ? 16 : BLOCK_N, p.f16 ? 0u : 1u);
        const uint64_t desc0 = make_sdesc(0);
        int s = 0;
        uint32_t ph = 0u;
        if (elect_one())
        for (int i = 0; i < nk; ++i) {
            mbar_wait(full_bar(s), ph);
            {
                const uint32_t st = smem_base + s * Cfg::STAGE;
                const uint64_t a_hi = desc0 + (uint64_t)((st & 0x3FFFF) >> 4), a_lo = a_hi + (uint64_t)(Cfg::A_TILE >> 4);
                const uint64_t w_hi = a_hi + (uint64_t)((Cfg::NA * Cfg::A_TILE) >> 4), w_lo = w_hi + (uint64_t)(Cfg::W_TILE >> 4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(tmem_acc, a_hi + 2 * kk, w_hi + 2 * kk, idesc, (i | kk) ? 1u : 0u);
                if (NPASS >= 2) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(tmem_acc, a_lo + 2 * kk, w_hi + 2 * kk, idesc, 1u);
                }
                if (NPASS >= 3) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(tmem_acc, a_hi + 2 * kk, w_lo + 2 * kk, idesc, 1u);
                }
                tc_commit(empty_bar(s));            // frees this smem stage when the MMAs above retire
                if (i == nk - 1) tc_commit(tmem_full_bar);       // accumulator complete -> epilogue
            }
            if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int tw = row % p.TW, th = (row / p.TW) % p.TH, tn = row / (p.TW * p.TH);
        const int n = n0 + tn, h = h0 + th, w = w0 + tw;
        const bool valid = (n < p.N) && (h < p.Ht) && (w < p.Wt);
        float* orow = out + ((((long)n * p.Ho + (long)h * p.os + p.ph + pa) * p.Wo) + (long)w * p.os + p.pw + pb) * p.Cout;
        if (nk > 0) {
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
        }
        const bool add_bias = (bias != nullptr) && (blockIdx.z == 0);
        const bool split = gridDim.z > 1;
        constexpr int CH = BLOCK_N >= 32 ? 32 : 16;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += CH) {
            uint32_t v[32];
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            if (nk > 0) {
                if (CH == 32) tc_ld32(taddr, v); else tc_ld16(taddr, v);
                tc_wait_ld();
            } else {
#pragma unroll
                for (int j = 0; j < CH; ++j) v[j] = 0u;
            }
            if (valid) {
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int co = co0 + c0 + j;
                    if (co < p.Cout) {
                        float f = __uint_as_float(v[j]) * p.out_scale;
                        if (add_bias) f += __ldg(bias + co);
                        if (split) atomicAdd(orow + co, f);
                        else {
                            if (p.act == DSR_ACT_TANH) f = tanhf(f);
                            v[j] = __float_as_uint(f);
                        }
                    }
                }
                if (!split) {
                    if ((p.Cout & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < CH; j += 4) {
                            const int co = co0 + c0 + j;
                            if (co < p.Cout)
                                *reinterpret_cast<uint4*>(orow + co) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CH; ++j) {
                            const int co = co0 + c0 + j;
                            if (co < p.Cout) orow[co] = __uint_as_float(v[j]);
                        }
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

// ------------------------------------------------------------------------------------------------
// weight-gradient GEMM:  dWp[cm][t*Ca + c] (+)= sum_{pixels p of the base grid} M[p + moff][cm] * A[p + aoff + tap_t][c]
//   M  : 16-bit arranged tensor whose channels index the rows of the weight gradient (dY for Conv2d, x for
//        ConvTranspose2d), A : the arranged tensor the forward conv multiplied with (x, or dY for ConvTranspose2d).
//   Both operands are read with the PIXEL dimension as the MMA K dimension, i.e. MN-major smem tiles: one TMA box
//   = 64 pixels x 64 channels (128-byte rows, SWIZZLE_128B); descriptor SBO = 1024 (next 8 pixels), LBO = 8192
//   (next 64 channels).  One CTA = one (tap, 128 x BLOCK_N) tile of dWp over a range of pixel tiles (split-K).
// ------------------------------------------------------------------------------------------------
struct WgParams {
    int N, Hb, Wb;            // base grid
    int TW, TH, TN;           // 64-pixel K tile
    int tiles_w, tiles_h, tiles_total;
    int tiles_per_split;
    int mh, mw;               // M-side coordinate offsets
    int ah, aw;               // A-side coordinate offsets
    int Cm_real, Ca, T, n_tiles_c;   // rows of dWp, arranged A channels, taps, number of BLOCK_N tiles over Ca
    int f16;
    float out_scale;
    signed char dr[TC_MAX_TAPS], ds[TC_MAX_TAPS];
};

__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int BLOCK_N, int NPASS>
struct WgCfg {
    static constexpr int BOX = 64 * 128;                     // 64 pixels x 128 B
    static constexpr int NB = BLOCK_N / 64;
    static constexpr int NM = NPASS >= 2 ? 2 : 1;            // M hi (+lo)
    static constexpr int NA = NPASS >= 3 ? 2 : 1;            // A hi (+lo)
    static constexpr int M_BYTES = 2 * BOX;                  // 128 channels
    static constexpr int A_BYTES = NB * BOX;
    static constexpr int STAGE = NM * M_BYTES + NA * A_BYTES;
    static constexpr int STAGES = (200 * 1024 / STAGE) > 8 ? 8 : (200 * 1024 / STAGE);
    static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
};

template <int BLOCK_N, int NPASS>
__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapM_hi, const __grid_constant__ CUtensorMap mapM_lo,
                const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                const __grid_constant__ WgParams p, float* __restrict__ dWp) {
    using Cfg = WgCfg<BLOCK_N, NPASS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * Cfg::STAGES);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * Cfg::STAGES + 1);
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::STAGES * Cfg::STAGE + 8 * (2 * Cfg::STAGES + 1));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int t = blockIdx.x;                                 // tap
    const int cm0 = (blockIdx.y / p.n_tiles_c) * 128;         // rows of dWp
    const int cn0 = (blockIdx.y % p.n_tiles_c) * BLOCK_N;     // arranged A channel
    const int kb = blockIdx.z * p.tiles_per_split;
    int ke = kb + p.tiles_per_split;
    if (ke > p.tiles_total) ke = p.tiles_total;
    const int nk = ke - kb;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapM_hi); tma_prefetch_desc(&mapA_hi);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)BLOCK_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_ptr_gen;

    if (warp == 0) {
        // TMA producer (whole warp loops and waits, one elected lane issues)
        for (int i = 0; i < nk; ++i) {
            const int s = i % Cfg::STAGES, it = i / Cfg::STAGES;
            mbar_wait(empty_bar(s), (it & 1) ^ 1);
            if (elect_one()) {
                int tile = kb + i;
                const int tw_i = tile % p.tiles_w; tile /= p.tiles_w;
                const int th_i = tile % p.tiles_h; tile /= p.tiles_h;
                const int n0 = tile * p.TN, h0 = th_i * p.TH, w0 = tw_i * p.TW;
                const uint32_t st = smem_base + s * Cfg::STAGE;
                mbar_expect_tx(full_bar(s), Cfg::STAGE);
                uint32_t dst = st;
#pragma unroll
                for (int b = 0; b < 2; ++b, dst += Cfg::BOX)
                    tma_load_4d(dst, &mapM_hi, full_bar(s), cm0 + 64 * b, w0 + p.mw, h0 + p.mh, n0);
                if (NPASS >= 2) {
#pragma unroll
                    for (int b = 0; b < 2; ++b, dst += Cfg::BOX)
                        tma_load_4d(dst, &mapM_lo, full_bar(s), cm0 + 64 * b, w0 + p.mw, h0 + p.mh, n0);
                }
                const int ha = h0 + p.ah + p.dr[t], wa = w0 + p.aw + p.ds[t];
#pragma unroll
                for (int b = 0; b < Cfg::NB; ++b, dst += Cfg::BOX)
                    tma_load_4d(dst, &mapA_hi, full_bar(s), cn0 + 64 * b, wa, ha, n0);
                if (NPASS >= 3) {
#pragma unroll
                    for (int b = 0; b < Cfg::NB; ++b, dst += Cfg::BOX)
                        tma_load_4d(dst, &mapA_lo, full_bar(s), cn0 + 64 * b, wa, ha, n0);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // MMA issuer; both operands MN-major: bits 15 / 16 of the instruction descriptor
        // operand formats are independent fields of the instruction descriptor (bits 7..9: A = M tensor, 10..12: B = A
        // tensor): dY is bf16 (gradients span too many octaves for half) while x may be the forward pass's f16 operand
        const uint32_t idesc = (1u << 4) | (((p.f16 & 1) ? 0u : 1u) << 7) | (((p.f16 & 2) ? 0u : 1u) << 10) |
                               ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) | (1u << 15) | (1u << 16);
        const uint64_t desc0 = make_sdesc_mn(0);
        for (int i = 0; i < nk; ++i) {
            const int s = i % Cfg::STAGES, it = i / Cfg::STAGES;
            mbar_wait(full_bar(s), it & 1);
            if (elect_one()) {
                const uint32_t st = smem_base + s * Cfg::STAGE;
                const uint64_t m_hi = desc0 + (uint64_t)((st & 0x3FFFF) >> 4), m_lo = m_hi + (uint64_t)(Cfg::M_BYTES >> 4);
                const uint64_t a_hi = m_hi + (uint64_t)((Cfg::NM * Cfg::M_BYTES) >> 4), a_lo = a_hi + (uint64_t)(Cfg::A_BYTES >> 4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(tmem_acc, m_hi + 128 * kk, a_hi + 128 * kk, idesc, (i | kk) ? 1u : 0u);
                if (NPASS >= 2) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(tmem_acc, m_lo + 128 * kk, a_hi + 128 * kk, idesc, 1u);
                }
                if (NPASS >= 3) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(tmem_acc, m_hi + 128 * kk, a_lo + 128 * kk, idesc, 1u);
                }
                tc_commit(empty_bar(s));
                if (i == nk - 1) tc_commit(tmem_full_bar);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int cm = cm0 + row;
        const bool valid = cm < p.Cm_real && nk > 0;
        float* orow = dWp + (long)cm * ((long)p.T * p.Ca) + (long)t * p.Ca;
        if (nk > 0) {
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
        }
        const bool split = gridDim.z > 1;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            uint32_t v[32];
            if (nk > 0) {
                tc_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                tc_wait_ld();
            }
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int c = cn0 + c0 + j;
                    if (c < p.Ca) {
                        float4 f = make_float4(__uint_as_float(v[j]) * p.out_scale, __uint_as_float(v[j + 1]) * p.out_scale,
                                               __uint_as_float(v[j + 2]) * p.out_scale, __uint_as_float(v[j + 3]) * p.out_scale);
                        if (split) {
                            atomicAdd(orow + c, f.x); atomicAdd(orow + c + 1, f.y);
                            atomicAdd(orow + c + 2, f.z); atomicAdd(orow + c + 3, f.w);
                        } else {
                            *reinterpret_cast<float4*>(orow + c) = f;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)BLOCK_N) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// operand preparation: fp32 NHWC -> arranged bf16 hi (+lo), fusing norm-apply + activation + padding
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int prep_pad_src(int q, int p, int n, int mode) {
    int i = q - p;
    if (i >= 0 && i < n) return i;
    if (mode == DSR_PAD_ZERO) return -1;
    if (mode == DSR_PAD_REFLECT) { i = i < 0 ? -i : 2 * (n - 1) - i; return (i >= 0 && i < n) ? i : -1; }
    return i < 0 ? 0 : n - 1;
}
// x = hi + lo with hi, lo in the 16-bit operand format (bf16: 8+8 significand bits; f16: 11+11, the low
// part degrading gracefully into half subnormals, i.e. an absolute error floor of 2^-25)
__device__ __forceinline__ void split16(float x, int f16, unsigned short& hi, unsigned short& lo) {
    if (f16) {
        x = fminf(fmaxf(x, -65504.f), 65504.f);
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        hi = __half_as_ushort(h); lo = __half_as_ushort(l);
    } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
        hi = __bfloat16_as_ushort(h); lo = __bfloat16_as_ushort(l);
    }
}
// two values -> packed 16-bit (hi, lo) pairs with the paired conversion instructions (cvt.rn.f16x2.f32 / bf16x2)
__device__ __forceinline__ void split16x2(float x0, float x1, int f16, unsigned& hi, unsigned& lo) {
    if (f16) {
        x0 = fminf(fmaxf(x0, -65504.f), 65504.f); x1 = fminf(fmaxf(x1, -65504.f), 65504.f);
        const __half2 h = __floats2half2_rn(x0, x1);
        const float2 b = __half22float2(h);
        const __half2 l = __floats2half2_rn(x0 - b.x, x1 - b.y);
        hi = *reinterpret_cast<const unsigned*>(&h); lo = *reinterpret_cast<const unsigned*>(&l);
    } else {
        const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
        const float2 b = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - b.x, x1 - b.y);
        hi = *reinterpret_cast<const unsigned*>(&h); lo = *reinterpret_cast<const unsigned*>(&l);
    }
}
// One block walks arranged rows (n, ha); one thread-item = 8 consecutive arranged channels of one arranged pixel (two
// 16-byte loads, one 16-byte store per output plane).  The kernel is instruction-bound, not memory-bound, so the item
// decomposition uses multiply-high "magic" divisions (exact for the item counts involved: it * d < 2^32), the
// (mean, scale, shift) table of the fused norm-apply is staged in shared memory per image and read with 16-byte loads,
// and the hi/lo split uses the paired conversion instructions.  csum (optional) receives the per-source-channel sum of
// everything written (the bias gradient when the tensor is dY), accumulated in shared memory per block and flushed with
// fp64 atomics.
// torch.cat(dim=1) folded into the preparation: up to 4 NHWC sources side by side along the channel axis (n = 0: `x` alone)
#define PREP_CAT_MAX_GROUPS 256      // concatenations of up to 2048 channels
struct PrepCat {
    const float* p[4];
    int c[4];            // channels of each source
    int n;
};
__device__ __forceinline__ const float* prep_cat_src(const PrepCat& cat, long pix, int c, int& room) {
    int off = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < cat.n && c < off + cat.c[k]) { room = off + cat.c[k] - c; return cat.p[k] + pix * cat.c[k] + (c - off); }
        off += k < cat.n ? cat.c[k] : 0;
    }
    room = 0;
    return nullptr;
}

__global__ void __launch_bounds__(256, 6)
tc_prep_kernel(const float* __restrict__ x, const PrepCat cat, int N, int H, int W, int C, const float* __restrict__ prm,
               int act, float slope, int pad, int mode, int layout, int Cp,
               unsigned short* __restrict__ Ahi, unsigned short* __restrict__ Alo, unsigned short* __restrict__ Abf,
               int Ha, int Wa, int Ca,
               int f16, double* __restrict__ csum, int csum_reps, unsigned magic_cg, unsigned magic_cp, unsigned magic_ha, const NormFin fin) {
    extern __shared__ __align__(16) float prep_sm[];
    const int Cs = (C + 3) & ~3;                  // 16-byte aligned parameter rows
    const bool has_prm = prm != nullptr || fin.sums != nullptr;
    float* s_prm = prep_sm;
    float* s_sum = prep_sm + (has_prm ? 3 * Cs : 0);
    const int tid = threadIdx.x;
    // concatenated sources: per 8-channel group of the concatenation, where it lives (built once per block) - base pointer of
    // its first channel, the source's pixel pitch, and whether the whole group sits 16-byte aligned inside ONE source
    __shared__ const float* s_cat_ptr[PREP_CAT_MAX_GROUPS];
    __shared__ int s_cat_pitch[PREP_CAT_MAX_GROUPS];
    if (cat.n) {
        for (int g8 = tid; g8 < (C + 7) / 8; g8 += 256) {
            int room;
            const float* q = prep_cat_src(cat, 0, g8 * 8, room);
            int pitch = 0, off = 0;
            for (int k = 0; k < cat.n; ++k) { if (g8 * 8 < off + cat.c[k]) { pitch = cat.c[k]; break; } off += cat.c[k]; }
            const bool whole = room >= 8 && (pitch & 3) == 0 && (((uintptr_t)q) & 15) == 0;
            s_cat_ptr[g8] = q;
            s_cat_pitch[g8] = whole ? pitch : -pitch;          // negative: element-wise path
        }
        __syncthreads();
    }
    const int cg = Ca >> 3, items = Wa * cg;
    const int Hq = H + 2 * pad, Wq = W + 2 * pad;
    const bool vec = (C & 3) == 0;
    const long NC = (long)N * C;
    if (csum) {
        for (int i = tid; i < C; i += 256) s_sum[i] = 0.f;
        __syncthreads();
    }
    // bias-gradient sums: when every item of a thread covers the same 8 channels (256 % cg == 0, NORMAL / S2D layout) they are
    // kept in registers and folded once at the end - the per-item shared-memory atomics were 16-way conflicted
    const bool reg_sum = csum && layout != DSR_TC_LAYOUT_PAIR && (256 % cg) == 0;
    float racc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int cached_n = -1;
    for (int row = blockIdx.x; row < N * Ha; row += gridDim.x) {
        const int n = magic_ha ? (int)__umulhi((unsigned)row, magic_ha) : row, ha = row - n * Ha;
        if (has_prm && n != cached_n) {
            __syncthreads();
            if (fin.sums) {
                // folded dsr_norm_finalize: every block derives the constants of its sample from the raw sums; the block that
                // owns the sample's first arranged row also writes them out for the readers of the backward pass
                for (int i = tid; i < C; i += 256) {
                    float m, sc, sh;
                    norm_fin_one(fin.sums, fin.P, fin.groups, fin.gamma, fin.beta, fin.eps, n, i, C, m, sc, sh);
                    s_prm[i] = m; s_prm[Cs + i] = sc; s_prm[2 * Cs + i] = sh;
                    if (ha == 0 && fin.prm_out) {
                        fin.prm_out[(long)n * C + i] = m; fin.prm_out[NC + (long)n * C + i] = sc; fin.prm_out[2 * NC + (long)n * C + i] = sh;
                    }
                }
            } else {
                for (int i = tid; i < C; i += 256) {
                    s_prm[i] = prm[(long)n * C + i]; s_prm[Cs + i] = prm[NC + (long)n * C + i]; s_prm[2 * Cs + i] = prm[2 * NC + (long)n * C + i];
                }
            }
            cached_n = n;
            __syncthreads();
        }
        for (int it = tid; it < items; it += 256) {
            const int wa = magic_cg ? (int)__umulhi((unsigned)it, magic_cg) : it, q = (it - wa * cg) << 3;
            int c, qi, qj;       // source channel start, padded-space row / col
            if (layout == DSR_TC_LAYOUT_NORMAL) { c = q; qi = ha; qj = wa; }
            else {
                const int g = magic_cp ? (int)__umulhi((unsigned)q, magic_cp) : q;
                c = q - g * Cp;
                if (layout == DSR_TC_LAYOUT_PAIR) { qi = ha; qj = wa + g; }
                else { qi = 2 * ha + (g >> 1); qj = 2 * wa + (g & 1); }
            }
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
            if (qi < Hq && qj < Wq && c < C) {
                const int i = prep_pad_src(qi, pad, H, mode), j = prep_pad_src(qj, pad, W, mode);
                if (i >= 0 && j >= 0) {
                    const long pix = (long)((n * H + i) * W + j);
                    const float* src;
                    bool full;
                    if (cat.n == 0) {
                        src = x + pix * C + c;
                        full = vec && c + 8 <= C;
                    } else {
                        const int pitch = s_cat_pitch[c >> 3];
                        full = pitch > 0;                                        // the 8 channels sit in ONE source, 16-byte aligned
                        src = s_cat_ptr[c >> 3] + pix * pitch;
                        if (!full) {                                             // a group that straddles two sources: element-wise
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (c + e < C) { int r2; const float* q2 = prep_cat_src(cat, pix, c + e, r2); v[e] = *q2; }
                        }
                    }
                    if (full) {
                        const float4 a = ld4(src), b4 = ld4(src + 4);
                        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b4.x; v[5] = b4.y; v[6] = b4.z; v[7] = b4.w;
                    } else if (cat.n == 0) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) if (c + e < C) v[e] = src[e];
                    }
                    if (has_prm) {
                        if (full) {
                            float m[8], sc[8], sh[8];
                            *reinterpret_cast<float4*>(m) = *reinterpret_cast<const float4*>(s_prm + c);
                            *reinterpret_cast<float4*>(m + 4) = *reinterpret_cast<const float4*>(s_prm + c + 4);
                            *reinterpret_cast<float4*>(sc) = *reinterpret_cast<const float4*>(s_prm + Cs + c);
                            *reinterpret_cast<float4*>(sc + 4) = *reinterpret_cast<const float4*>(s_prm + Cs + c + 4);
                            *reinterpret_cast<float4*>(sh) = *reinterpret_cast<const float4*>(s_prm + 2 * Cs + c);
                            *reinterpret_cast<float4*>(sh + 4) = *reinterpret_cast<const float4*>(s_prm + 2 * Cs + c + 4);
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = (v[e] - m[e]) * sc[e] + sh[e];
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (c + e < C) v[e] = (v[e] - s_prm[c + e]) * s_prm[Cs + c + e] + s_prm[2 * Cs + c + e];
                        }
                    }
                    if (act == DSR_ACT_RELU) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
                    } else if (act == DSR_ACT_LRELU) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = v[e] > 0.f ? v[e] : slope * v[e];
                    }
                    if (reg_sum) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) racc[e] += v[e];
                    } else if (csum) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) if (c + e < C) atomicAdd(&s_sum[c + e], v[e]);
                    }
                }
            }
            uint4 hi, lo;
            split16x2(v[0], v[1], f16, hi.x, lo.x); split16x2(v[2], v[3], f16, hi.y, lo.y);
            split16x2(v[4], v[5], f16, hi.z, lo.z); split16x2(v[6], v[7], f16, hi.w, lo.w);
            const long o = ((long)row * Wa + wa) * Ca + q;
            *reinterpret_cast<uint4*>(Ahi + o) = hi;
            if (Alo) *reinterpret_cast<uint4*>(Alo + o) = lo;
            if (Abf) {       // the same operand once more as plain bf16: what the weight-gradient GEMM of the backward pass reads
                uint4 b;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
                b.x = *reinterpret_cast<unsigned*>(&t0); b.y = *reinterpret_cast<unsigned*>(&t1);
                b.z = *reinterpret_cast<unsigned*>(&t2); b.w = *reinterpret_cast<unsigned*>(&t3);
                *reinterpret_cast<uint4*>(Abf + o) = b;
            }
        }
    }
    if (csum) {
        if (reg_sum) {
            const int q0 = (tid % cg) << 3;
            const int c = layout == DSR_TC_LAYOUT_NORMAL ? q0 : q0 - (q0 / Cp) * Cp;
#pragma unroll
            for (int e = 0; e < 8; ++e) if (c + e < C) atomicAdd(&s_sum[c + e], racc[e]);
        }
        __syncthreads();
        // csum_reps replicas of the accumulator row: block b adds into replica b % reps, so a launch of G blocks puts G / reps
        // (not G) serialised fp64 atomics on each address and the grid can stay wide enough to saturate HBM
        double* dst = csum + (long)(blockIdx.x % csum_reps) * C;
        for (int i = tid; i < C; i += 256) atomicAdd(&dst[i], (double)s_sum[i]);
    }
}

// The same preparation for the shapes the step is made of (no concatenation, C % 8 == 0, Ca / 8 a power of two <= 256): a
// thread keeps ONE group of 8 arranged channels for the whole launch, so everything that depends on the channel group -
// the layout's pixel offset, the source channel, the (mean, scale, shift) constants of the fused norm-apply, the bias-gradient
// partial sums - lives in registers, the arranged pixel advances by a constant step (no index divisions) and two items are
// in flight per thread.  Bit-identical to tc_prep_kernel (tests); DSR_PREP_FAST=0 keeps the general kernel for A/B.
// YOUT: the pass also IS the normalisation layer of a residual block - y = act(norm(x)) + res leaves as fp32 (written once per
// source pixel, from its un-padded position) next to the arranged operand of the next convolution, which is made from y
// (dsr_tc_prep_norm_res: one pass instead of dsr_norm_apply_fwd followed by dsr_tc_prep of its output).
template <bool PRM, bool CSUM, bool YOUT>
__global__ void __launch_bounds__(256, (PRM ? 3 : 4))
tc_prep_fast_kernel(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y_out,
                    int N, int H, int W, int C, const float* __restrict__ prm, int act, float slope,
                    int pad, int mode, int layout, int Cp, unsigned short* __restrict__ Ahi, unsigned short* __restrict__ Alo,
                    unsigned short* __restrict__ Abf, int Ha, int Wa, int Ca, int f16, double* __restrict__ csum, int csum_reps,
                    unsigned magic_ha, const NormFin fin) {
    extern __shared__ __align__(16) float prep_sm[];
    const int Cs = (C + 3) & ~3;
    float* s_prm = prep_sm;
    float* s_sum = prep_sm + (PRM ? 3 * Cs : 0);
    const int tid = threadIdx.x;
    const int cg = Ca >> 3, step = 256 / cg;             // 256 % cg == 0 (host)
    const int q = (tid & (cg - 1)) << 3, wa0 = tid / cg;
    // layout: arranged pixel (ha, wa), channel group q  ->  padded-space pixel (ha * mh + gh, wa * mw + gw), source channel c
    int c = q, mh = 1, mw = 1, gh = 0, gw = 0;
    if (layout != DSR_TC_LAYOUT_NORMAL) {
        const int g = q / Cp;
        c = q - g * Cp;
        if (layout == DSR_TC_LAYOUT_PAIR) gw = g;
        else { mh = 2; mw = 2; gh = g >> 1; gw = g & 1; }
    }
    const bool chan_ok = c < C;                           // (c + 8 <= C then: C % 8 == 0)
    const int Hq = H + 2 * pad, Wq = W + 2 * pad;
    const long NC = (long)N * C;
    if (CSUM) {
        for (int i = tid; i < C; i += 256) s_sum[i] = 0.f;
        __syncthreads();
    }
    float racc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float4 m0, m1, sc0, sc1, sh0, sh1;
    m0 = m1 = sc0 = sc1 = sh0 = sh1 = make_float4(0.f, 0.f, 0.f, 0.f);
    int cached_n = -1;
    for (int row = blockIdx.x; row < N * Ha; row += gridDim.x) {
        const int n = magic_ha ? (int)__umulhi((unsigned)row, magic_ha) : row, ha = row - n * Ha;
        if (PRM && n != cached_n) {
            __syncthreads();
            if (fin.sums) {
                for (int i = tid; i < C; i += 256) {
                    float m, sc, sh;
                    norm_fin_one(fin.sums, fin.P, fin.groups, fin.gamma, fin.beta, fin.eps, n, i, C, m, sc, sh);
                    s_prm[i] = m; s_prm[Cs + i] = sc; s_prm[2 * Cs + i] = sh;
                    if (ha == 0 && fin.prm_out) {
                        fin.prm_out[(long)n * C + i] = m; fin.prm_out[NC + (long)n * C + i] = sc; fin.prm_out[2 * NC + (long)n * C + i] = sh;
                    }
                }
            } else {
                for (int i = tid; i < C; i += 256) {
                    s_prm[i] = prm[(long)n * C + i]; s_prm[Cs + i] = prm[NC + (long)n * C + i]; s_prm[2 * Cs + i] = prm[2 * NC + (long)n * C + i];
                }
            }
            cached_n = n;
            __syncthreads();
            if (chan_ok) {
                m0 = ld4(s_prm + c); m1 = ld4(s_prm + c + 4);
                sc0 = ld4(s_prm + Cs + c); sc1 = ld4(s_prm + Cs + c + 4);
                sh0 = ld4(s_prm + 2 * Cs + c); sh1 = ld4(s_prm + 2 * Cs + c + 4);
            }
        }
        const int qi = ha * mh + gh;
        const int i = (qi < Hq && chan_ok) ? prep_pad_src(qi, pad, H, mode) : -1;
        const float* xrow = x + ((long)(n * H + (i >= 0 ? i : 0)) * W) * C + c;
        const long obase = (long)row * Wa * Ca + q;
        for (int wa = wa0; wa < Wa; wa += 2 * step) {
            // two items per thread and iteration: both loads are issued before either is consumed
            const int wb = wa + step;
            const int qja = wa * mw + gw, qjb = wb * mw + gw;
            const int ja = (i >= 0 && qja < Wq) ? prep_pad_src(qja, pad, W, mode) : -1;
            const int jb = (i >= 0 && wb < Wa && qjb < Wq) ? prep_pad_src(qjb, pad, W, mode) : -1;
            float4 a0, a1, b0, b1, ra0, ra1, rb0, rb1;
            a0 = a1 = b0 = b1 = ra0 = ra1 = rb0 = rb1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ja >= 0) { a0 = ld4(xrow + (long)ja * C); a1 = ld4(xrow + (long)ja * C + 4); }
            if (jb >= 0) { b0 = ld4(xrow + (long)jb * C); b1 = ld4(xrow + (long)jb * C + 4); }
            if (YOUT && res) {
                const float* rrow = res + (xrow - x);
                if (ja >= 0) { ra0 = ld4(rrow + (long)ja * C); ra1 = ld4(rrow + (long)ja * C + 4); }
                if (jb >= 0) { rb0 = ld4(rrow + (long)jb * C); rb1 = ld4(rrow + (long)jb * C + 4); }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && wb >= Wa) break;
                const bool valid = u == 0 ? ja >= 0 : jb >= 0;
                const float4 p0 = u == 0 ? a0 : b0, p1 = u == 0 ? a1 : b1;
                float v[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                if (valid) {
                    if (PRM) {
                        v[0] = (v[0] - m0.x) * sc0.x + sh0.x; v[1] = (v[1] - m0.y) * sc0.y + sh0.y;
                        v[2] = (v[2] - m0.z) * sc0.z + sh0.z; v[3] = (v[3] - m0.w) * sc0.w + sh0.w;
                        v[4] = (v[4] - m1.x) * sc1.x + sh1.x; v[5] = (v[5] - m1.y) * sc1.y + sh1.y;
                        v[6] = (v[6] - m1.z) * sc1.z + sh1.z; v[7] = (v[7] - m1.w) * sc1.w + sh1.w;
                    }
                    if (act == DSR_ACT_RELU) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
                    } else if (act == DSR_ACT_LRELU) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = v[e] > 0.f ? v[e] : slope * v[e];
                    }
                    if (CSUM) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) racc[e] += v[e];
                    }
                    if (YOUT) {
                        const float4 r0 = u == 0 ? ra0 : rb0, r1 = u == 0 ? ra1 : rb1;
                        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
                        const int jj = u == 0 ? ja : jb, qj = u == 0 ? qja : qjb;
                        if (qi - pad == i && qj - pad == jj) {               // the pixel's own (un-padded) position writes y
                            float* yp = y_out + (xrow - x) + (long)jj * C;
                            st4(yp, make_float4(v[0], v[1], v[2], v[3])); st4(yp + 4, make_float4(v[4], v[5], v[6], v[7]));
                        }
                    }
                }
                uint4 hi, lo;
                split16x2(v[0], v[1], f16, hi.x, lo.x); split16x2(v[2], v[3], f16, hi.y, lo.y);
                split16x2(v[4], v[5], f16, hi.z, lo.z); split16x2(v[6], v[7], f16, hi.w, lo.w);
                const long o = obase + (long)(u == 0 ? wa : wb) * Ca;
                *reinterpret_cast<uint4*>(Ahi + o) = hi;
                if (Alo) *reinterpret_cast<uint4*>(Alo + o) = lo;
                if (Abf) {
                    uint4 b;
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
                    __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
                    b.x = *reinterpret_cast<unsigned*>(&t0); b.y = *reinterpret_cast<unsigned*>(&t1);
                    b.z = *reinterpret_cast<unsigned*>(&t2); b.w = *reinterpret_cast<unsigned*>(&t3);
                    *reinterpret_cast<uint4*>(Abf + o) = b;
                }
            }
        }
    }
    if (CSUM) {
        if (chan_ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&s_sum[c + e], racc[e]);
        }
        __syncthreads();
        double* dst = csum + (long)(blockIdx.x % csum_reps) * C;
        for (int i = tid; i < C; i += 256) atomicAdd(&dst[i], (double)s_sum[i]);
    }
}

// InstanceNorm2d(affine=False) [+ReLU] backward AND the operand of the NEXT backward GEMM in one pass: the gradient that
// leaves a normalisation layer is the dY of the convolution in front of it, whose data- and weight-gradient GEMMs read it as a
// zero-padded arranged 16-bit tensor.  One thread = one group of 8 arranged channels (the scheme of tc_prep_fast_kernel):
//   dx = rstd * (g' - mean(g') - xhat * mean(g' * xhat))   (fp32 NHWC, what dsr_in_bwd_apply writes - same arithmetic, same bits)
//   A  = arranged hi (+lo) planes of dx with a zero frame of `pad` pixels (what dsr_tc_prep makes of dx afterwards)
//   csum += per-channel sums of dx (the bias gradient of that convolution), replicated rows as in dsr_tc_prep.
// Zero padding only: every source pixel then owns exactly one arranged position, so dx and the sums are written / counted once.
template <bool CSUM>
__global__ void __launch_bounds__(256, 2)
tc_prep_inbwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, int N, int H, int W, int C,
                     const float* __restrict__ prm, const double* __restrict__ sums2, int act, int pad, int layout, int Cp,
                     unsigned short* __restrict__ Ahi, unsigned short* __restrict__ Alo, int Ha, int Wa, int Ca, int f16,
                     double* __restrict__ csum, int csum_reps, unsigned magic_ha) {
    extern __shared__ __align__(16) float prep_sm[];
    const int Cs = (C + 3) & ~3;
    float* s_c = prep_sm;                                 // [4][Cs]: mean, rstd, mean(g'), mean(g' * xhat)
    float* s_sum = prep_sm + 4 * Cs;
    const int tid = threadIdx.x;
    const int cg = Ca >> 3, step = 256 / cg;
    const int q = (tid & (cg - 1)) << 3, wa0 = tid / cg;
    int c = q, mh = 1, mw = 1, gh = 0, gw = 0;
    if (layout != DSR_TC_LAYOUT_NORMAL) {                 // S2D: arranged pixel (ha, wa), group g -> padded pixel (2 ha + g / 2, 2 wa + g % 2)
        const int g = q / Cp;
        c = q - g * Cp;
        mh = 2; mw = 2; gh = g >> 1; gw = g & 1;
    }
    const bool chan_ok = c < C;
    const bool relu = act == DSR_ACT_RELU;
    const int Hq = H + 2 * pad, Wq = W + 2 * pad;
    const long NC = (long)N * C;
    const float invP = 1.f / (float)((long)H * W);
    if (CSUM) {
        for (int i = tid; i < C; i += 256) s_sum[i] = 0.f;
        __syncthreads();
    }
    float racc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float4 mu0, mu1, rs0, rs1, a10, a11, a20, a21;
    mu0 = mu1 = rs0 = rs1 = a10 = a11 = a20 = a21 = make_float4(0.f, 0.f, 0.f, 0.f);
    int cached_n = -1;
    for (int row = blockIdx.x; row < N * Ha; row += gridDim.x) {
        const int n = magic_ha ? (int)__umulhi((unsigned)row, magic_ha) : row, ha = row - n * Ha;
        if (n != cached_n) {
            __syncthreads();
            for (int i = tid; i < C; i += 256) {
                s_c[i] = prm[(long)n * C + i];
                s_c[Cs + i] = prm[NC + (long)n * C + i];
                s_c[2 * Cs + i] = (float)sums2[((long)n * C + i) * 2] * invP;
                s_c[3 * Cs + i] = (float)sums2[((long)n * C + i) * 2 + 1] * invP;
            }
            cached_n = n;
            __syncthreads();
            if (chan_ok) {
                mu0 = ld4(s_c + c); mu1 = ld4(s_c + c + 4);
                rs0 = ld4(s_c + Cs + c); rs1 = ld4(s_c + Cs + c + 4);
                a10 = ld4(s_c + 2 * Cs + c); a11 = ld4(s_c + 2 * Cs + c + 4);
                a20 = ld4(s_c + 3 * Cs + c); a21 = ld4(s_c + 3 * Cs + c + 4);
            }
        }
        const int qi = ha * mh + gh;
        const int i = (qi < Hq && chan_ok) ? prep_pad_src(qi, pad, H, DSR_PAD_ZERO) : -1;
        const long srow = ((long)(n * H + (i >= 0 ? i : 0)) * W) * C + c;
        const long obase = (long)row * Wa * Ca + q;
        for (int wa = wa0; wa < Wa; wa += 2 * step) {
            const int wb = wa + step;
            const int qja = wa * mw + gw, qjb = wb * mw + gw;
            const int ja = (i >= 0 && qja < Wq) ? prep_pad_src(qja, pad, W, DSR_PAD_ZERO) : -1;
            const int jb = (i >= 0 && wb < Wa && qjb < Wq) ? prep_pad_src(qjb, pad, W, DSR_PAD_ZERO) : -1;
            float4 ga0, ga1, gb0, gb1, xa0, xa1, xb0, xb1;
            ga0 = ga1 = gb0 = gb1 = xa0 = xa1 = xb0 = xb1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ja >= 0) {
                const long o = srow + (long)ja * C;
                ga0 = ld4(dy + o); ga1 = ld4(dy + o + 4); xa0 = ld4(x + o); xa1 = ld4(x + o + 4);
            }
            if (jb >= 0) {
                const long o = srow + (long)jb * C;
                gb0 = ld4(dy + o); gb1 = ld4(dy + o + 4); xb0 = ld4(x + o); xb1 = ld4(x + o + 4);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && wb >= Wa) break;
                const int jj = u == 0 ? ja : jb;
                const float4 g0 = u == 0 ? ga0 : gb0, g1 = u == 0 ? ga1 : gb1, x0 = u == 0 ? xa0 : xb0, x1 = u == 0 ? xa1 : xb1;
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (jj >= 0) {
                    v[0] = in_bwd_one(x0.x, g0.x, mu0.x, rs0.x, a10.x, a20.x, relu);
                    v[1] = in_bwd_one(x0.y, g0.y, mu0.y, rs0.y, a10.y, a20.y, relu);
                    v[2] = in_bwd_one(x0.z, g0.z, mu0.z, rs0.z, a10.z, a20.z, relu);
                    v[3] = in_bwd_one(x0.w, g0.w, mu0.w, rs0.w, a10.w, a20.w, relu);
                    v[4] = in_bwd_one(x1.x, g1.x, mu1.x, rs1.x, a11.x, a21.x, relu);
                    v[5] = in_bwd_one(x1.y, g1.y, mu1.y, rs1.y, a11.y, a21.y, relu);
                    v[6] = in_bwd_one(x1.z, g1.z, mu1.z, rs1.z, a11.z, a21.z, relu);
                    v[7] = in_bwd_one(x1.w, g1.w, mu1.w, rs1.w, a11.w, a21.w, relu);
                    float* dp = dx + srow + (long)jj * C;
                    st4(dp, make_float4(v[0], v[1], v[2], v[3])); st4(dp + 4, make_float4(v[4], v[5], v[6], v[7]));
                    if (CSUM) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) racc[e] += v[e];
                    }
                }
                uint4 hi, lo;
                split16x2(v[0], v[1], f16, hi.x, lo.x); split16x2(v[2], v[3], f16, hi.y, lo.y);
                split16x2(v[4], v[5], f16, hi.z, lo.z); split16x2(v[6], v[7], f16, hi.w, lo.w);
                const long o = obase + (long)(u == 0 ? wa : wb) * Ca;
                *reinterpret_cast<uint4*>(Ahi + o) = hi;
                if (Alo) *reinterpret_cast<uint4*>(Alo + o) = lo;
            }
        }
    }
    if (CSUM) {
        if (chan_ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&s_sum[c + e], racc[e]);
        }
        __syncthreads();
        double* dst = csum + (long)(blockIdx.x % csum_reps) * C;
        for (int i = tid; i < C; i += 256) atomicAdd(&dst[i], (double)s_sum[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// weight packing: 4-D fp32 parameter -> bf16 hi (+lo) [Cout][T*Ca], K-major
//   variant CONV     : Conv2d weight (Cout, Cin, R, S); tap t = r*S + s, channel c          (stride 1)
//   variant CONV_PAIR: as CONV with Ca = g*Cp (g pixels side by side): tap t = r*ceil(S/g) + s/g, q = (s%g)*Cp + c
//   variant CONV_S2D : stride-2 conv (k <= 4) as 2x2 taps over 4*Cp channels: r = 2r'+a, s = 2s'+b
//   variant CONV_DGRAD: Conv2d weight (Cout, Cin, R, S) for the stride-1 data gradient: GEMM output channel = ci,
//                      K channel = co, tap t = (r', s') reads W[co][ci][R-1-r'][S-1-s']
//   variant CONV_DGRAD_PAIR: the same data gradient for narrow layers (Cin = 32): GEMM row co' = dx*Cin + ci is output
//                      channel ci of the pixel at column g*k + dx, the A operand is read as GROUPS of g = Ca / Cp adjacent pixels
//                      (pairs, or quads: 4 x 32 = 128 GEMM rows): tap t = (r', a), K channel q = b*Cp + co reads
//                      W[co][ci][R-1-r'][S-1-(g*a+b-dx)] (zero outside)
//   variant CONVT_PH : ConvTranspose2d weight (Cin, Cout, R, S), stride 2, phase (a,b): taps (dr,ds) in {0,1}^2,
//                      kh = pad + 2 - a - 2*dr, kw = pad + 2 - b - 2*ds (zero tap when outside the kernel)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_map_k(int variant, int t, int q, int R, int S, int Cp, int g, int pa, int pb, int pad,
                                         int& r, int& s, int& c) {
    if (variant == DSR_TC_W_CONV) { r = t / S; s = t - r * S; c = q; }
    else if (variant == DSR_TC_W_CONV_PAIR) { const int Sg = (S + g - 1) / g; r = t / Sg; s = g * (t - r * Sg) + q / Cp; c = q % Cp; }
    else if (variant == DSR_TC_W_CONV_S2D) { const int ab = q / Cp; c = q - ab * Cp; r = 2 * (t >> 1) + (ab >> 1); s = 2 * (t & 1) + (ab & 1); }
    else if (variant == DSR_TC_W_CONV_DGRAD) { r = R - 1 - t / S; s = S - 1 - (t - (t / S) * S); c = q; }
    else { c = q; r = pad + 2 - pa - 2 * (t >> 1); s = pad + 2 - pb - 2 * (t & 1); }
}
// one thread = 8 consecutive K positions of one output row (same tap, 8 consecutive arranged channels): eight gathered
// 4-byte loads (the parameter is small and L2-resident) and one 16-byte store per plane
__global__ void __launch_bounds__(256)
tc_pack_weight_kernel(const float* __restrict__ w, int D0, int D1, int R, int S, int variant, int Cp,
                      int pa, int pb, int pad, int Cout, int T, int Ca,
                      unsigned short* __restrict__ Whi, unsigned short* __restrict__ Wlo, int f16, float wscale) {
    const int Ca8 = Ca >> 3;
    const long total = (long)Cout * T * Ca8;
    if (pa < 0) {                 // all four output phases of a stride-2 transposed conv in one launch (grid.y = phase)
        pa = blockIdx.y >> 1; pb = blockIdx.y & 1;
        Whi += (long)blockIdx.y * Cout * T * Ca;
        if (Wlo) Wlo += (long)blockIdx.y * Cout * T * Ca;
    }
    const bool convT = (variant == DSR_TC_W_CONVT_PH) || (variant == DSR_TC_W_CONV_DGRAD);
    const int Cin = convT ? D0 : D1;
    const int g = Ca / Cp;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        // taps fastest: the R x S taps of one (co, c) pair are contiguous in the parameter, so a warp's gathered 4-byte
        // loads fall into a few sectors (channel-fastest order touched 32 sectors per load: the kernel was L2-gather bound)
        const int t = (int)(idx % T);
        const long u = idx / T;
        const int q8 = (int)(u % Ca8), co = (int)(u / Ca8);
        __align__(16) unsigned short hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int r, sx, c;
            float v = 0.f;
            if (variant == DSR_TC_W_CONV_DGRAD_PAIR) {
                // g = Ca / Cp adjacent pixels per GEMM row (2: pairs, N = 64; 4: quads, N = 128 = one channel-major M tile)
                const int q = q8 * 8 + e, Sg = (g + S - 2) / g + 1;
                const int dx = co / D1, ci = co - dx * D1, rr = t / Sg, a = t - rr * Sg, b = q / Cp, cc = q - b * Cp;
                const int sidx = g * a + b - dx;
                if (sidx >= 0 && sidx < S && rr < R && cc < D0 && dx < g)
                    v = w[(((long)cc * D1 + ci) * R + (R - 1 - rr)) * S + (S - 1 - sidx)];
                split16(v * wscale, f16, hi[e], lo[e]);
                continue;
            }
            tc_map_k(variant, t, q8 * 8 + e, R, S, Cp, g, pa, pb, pad, r, sx, c);
            // convT-style indexing: the GEMM's K channel is the parameter's dim 0, its output channel dim 1
            if (r >= 0 && r < R && sx >= 0 && sx < S && c >= 0 && c < Cin)
                v = convT ? w[(((long)c * D1 + co) * R + r) * S + sx] : w[(((long)co * D1 + c) * R + r) * S + sx];
            split16(v * wscale, f16, hi[e], lo[e]);
        }
        const long o = ((long)co * T + t) * Ca + q8 * 8;
        *reinterpret_cast<uint4*>(Whi + o) = *reinterpret_cast<const uint4*>(hi);
        if (Wlo) *reinterpret_cast<uint4*>(Wlo + o) = *reinterpret_cast<const uint4*>(lo);
    }
}
// Tiled form of the same packing for the big layers (R * S <= 16; variants CONV, CONV_S2D, CONV_DGRAD, CONVT_PH).  The gather
// kernel above reads coalesced but stores 16-byte pieces 2 * Ca bytes apart (a warp's store touches 32 lines), and it was
// the slowest of the support kernels: 1.4 ms per step for 61 launches over the 44 M trainable parameters (r2a) - ~0.5 TB/s.
// Here a CTA stages a tile of the parameter in shared memory with contiguous loads - TCO GEMM rows x TC K-channels x all R*S
// taps: for a Conv2d-style source (K channel = parameter dim 1) TC * R * S floats per row are one contiguous run, for a
// ConvTranspose-style source (K channel = dim 0) TCO * R * S floats per K channel are - and then writes whole
// [co][t][c0 .. c0 + TC) runs, 16 bytes per thread with the lanes side by side.  Bit-identical output (tests).
template <int VARIANT, int RS>
__global__ void __launch_bounds__(256)
tc_pack_weight_tiled_kernel(const float* __restrict__ w, int D0, int D1, int Cp, int pa, int pb, int pad, int Cout, int Ca,
                            unsigned short* __restrict__ Whi, unsigned short* __restrict__ Wlo, int f16, float wscale) {
    constexpr bool CONVT = VARIANT == DSR_TC_W_CONVT_PH || VARIANT == DSR_TC_W_CONV_DGRAD;
    constexpr int S = RS == 9 ? 3 : 4, R = S;
    constexpr int NAB = VARIANT == DSR_TC_W_CONV_S2D ? 4 : 1;
    constexpr int T = (VARIANT == DSR_TC_W_CONV_S2D || VARIANT == DSR_TC_W_CONVT_PH) ? 4 : RS;
    constexpr int TCO = CONVT ? 8 : 2, TC = CONVT ? 64 : 128, TCP = TC + 4, C8N = TC / 8;
    // shared tile [TCO][RS][TC + 4]: K channels fastest, so a thread's 8 output channels are two 16-byte loads and the lanes of
    // a warp read side by side (conflict-free); every index division below is by a compile-time constant
    __shared__ __align__(16) float wtile[TCO * RS * TCP];
    if (pa < 0) {                 // all four output phases of a stride-2 transposed conv in one launch (grid.y = phase)
        pa = blockIdx.y >> 1; pb = blockIdx.y & 1;
        Whi += (long)blockIdx.y * Cout * T * Ca;
        if (Wlo) Wlo += (long)blockIdx.y * Cout * T * Ca;
    }
    const int Cin = CONVT ? D0 : D1, Crow = CONVT ? D1 : D0;          // K channels / GEMM rows the parameter really has
    const int Kc = NAB == 4 ? Cp : Ca;                                // K-channel extent of one (tap, sub-block)
    const int nc_tiles = (Kc + TC - 1) / TC, nco_tiles = (Cout + TCO - 1) / TCO;
    const int tid = threadIdx.x;
    for (int tile = blockIdx.x; tile < nc_tiles * nco_tiles; tile += gridDim.x) {
        const int tco = tile / nc_tiles, co0 = tco * TCO, c0 = (tile - tco * nc_tiles) * TC;
        const int nc = min(TC, Cin - c0), nrow = min(TCO, Crow - co0);    // (<= 0: the tile lies in the zero padding)
        __syncthreads();
        if (nc > 0 && nrow > 0) {
            if (!CONVT) {
                for (int co_l = 0; co_l < nrow; ++co_l) {
                    const float* src = w + ((long)(co0 + co_l) * D1 + c0) * RS;
                    for (int i = tid; i < nc * RS; i += 256) { const int c_l = i / RS; wtile[(co_l * RS + (i - c_l * RS)) * TCP + c_l] = __ldg(src + i); }
                }
            } else {
                const int run = nrow * RS;                            // contiguous floats per K channel
                for (int c_l = tid >> 5; c_l < nc; c_l += 8) {        // one warp per K channel
                    const float* src = w + ((long)(c0 + c_l) * D1 + co0) * RS;
                    for (int k = tid & 31; k < run; k += 32) { const int co_l = k / RS; wtile[(co_l * RS + (k - co_l * RS)) * TCP + c_l] = __ldg(src + k); }
                }
            }
        }
        __syncthreads();
        for (int it = tid; it < TCO * T * NAB * C8N; it += 256) {
            const int c8_l = it % C8N; int u = it / C8N;
            const int ab = u % NAB; u /= NAB;
            const int t = u % T, co_l = u / T, co = co0 + co_l;
            const int cb = c0 + c8_l * 8, q0 = ab * Cp + cb;
            if (co >= Cout || cb >= Kc) continue;
            int r, sx, c;
            tc_map_k(VARIANT, t, q0, R, S, Cp, 1, pa, pb, pad, r, sx, c);
            const bool tap_ok = r >= 0 && r < R && sx >= 0 && sx < S && co_l < nrow && cb < Cin;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (tap_ok) {
                const float4* row = reinterpret_cast<const float4*>(wtile + (co_l * RS + r * S + sx) * TCP + c8_l * 8);
                const float4 a = row[0], b = row[1];
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            }
            uint4 hi, lo;
            unsigned h2[4], l2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float x0 = cb + 2 * e < Cin ? v[2 * e] * wscale : 0.f, x1 = cb + 2 * e + 1 < Cin ? v[2 * e + 1] * wscale : 0.f;
                split16x2(x0, x1, f16, h2[e], l2[e]);
            }
            hi = make_uint4(h2[0], h2[1], h2[2], h2[3]); lo = make_uint4(l2[0], l2[1], l2[2], l2[3]);
            const long o = ((long)co * T + t) * Ca + q0;
            *reinterpret_cast<uint4*>(Whi + o) = hi;
            if (Wlo) *reinterpret_cast<uint4*>(Wlo + o) = lo;
        }
    }
}
template <int VARIANT, int RS>
static void pack_tiled_launch(const float* w, int D0, int D1, int Cp, int pa, int pb, int pad, int Cout, int Ca, void* W_hi, void* W_lo,
                              int f16, float wscale, cudaStream_t st) {
    constexpr bool CONVT = VARIANT == DSR_TC_W_CONVT_PH || VARIANT == DSR_TC_W_CONV_DGRAD;
    constexpr int TCO = CONVT ? 8 : 2, TC = CONVT ? 64 : 128;
    const int Kc = VARIANT == DSR_TC_W_CONV_S2D ? Cp : Ca;
    const long tiles = (long)dsr_cdiv(Kc, TC) * dsr_cdiv(Cout, TCO);
    const long cap = (long)dsr_num_sms() * (pa < 0 ? 2 : 6);
    const dim3 grid((unsigned)(tiles < cap ? tiles : cap), pa < 0 ? 4 : 1);
    tc_pack_weight_tiled_kernel<VARIANT, RS><<<grid, 256, 0, st>>>(w, D0, D1, Cp, pa, pb, pad, Cout, Ca, (unsigned short*)W_hi,
                                                                  (unsigned short*)W_lo, f16, wscale);
}
// packed weight gradient [D0][T*Ca] (row = parameter dim 0, K channel = parameter dim 1) -> parameter layout
__global__ void tc_unpack_wgrad_kernel(const float* __restrict__ dWp, int D0, int D1, int R, int S, int variant, int Cp,
                                       int T, int Ca, float* __restrict__ grad, int accumulate, int nsplit, long split_stride) {
    const long total = (long)D0 * D1 * R * S;
    // gather form (one thread per parameter element): invert tc_map_k
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int s = (int)(idx % S); long u = idx / S;
        const int r = (int)(u % R); u /= R;
        const int c = (int)(u % D1); const int d0 = (int)(u / D1);
        int t, q;
        if (variant == DSR_TC_W_CONV) { t = r * S + s; q = c; }
        else if (variant == DSR_TC_W_CONV_PAIR) { const int g = Ca / Cp, Sg = (S + g - 1) / g; t = r * Sg + s / g; q = (s % g) * Cp + c; }
        else { t = (r >> 1) * 2 + (s >> 1); q = ((r & 1) * 2 + (s & 1)) * Cp + c; }          // S2D
        const float* src = dWp + (long)d0 * ((long)T * Ca) + (long)t * Ca + q;
        float v = src[0];
        for (int z = 1; z < nsplit; ++z) v += src[z * split_stride];      // K-split slabs of dsr_tc_wgrad2p, fixed order
        if (accumulate) grad[idx] += v; else grad[idx] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int NPASS>
static int launch_tc(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
                     const TcParams& p, const float* bias, float* out, dim3 grid, cudaStream_t st) {
    using Cfg = TcCfg<BLOCK_N, NPASS>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) {
            dsr_set_error("conv_tc: cannot raise dynamic shared memory to %d", Cfg::SMEM);
            return DSR_ERR_CUDA;
        }
        attr = true;
    }
    conv_tc_kernel<BLOCK_N, NPASS><<<grid, 192, Cfg::SMEM, st>>>(ah, al, wh, wl, p, bias, out);
    return dsr_check_launch("conv_tc");
}

template <int NPASS>
static int dispatch_n(int bn, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
                      const TcParams& p, const float* bias, float* out, dim3 grid, cudaStream_t st) {
    switch (bn) {
        case 16: return launch_tc<16, NPASS>(ah, al, wh, wl, p, bias, out, grid, st);
        case 32: return launch_tc<32, NPASS>(ah, al, wh, wl, p, bias, out, grid, st);
        case 64: return launch_tc<64, NPASS>(ah, al, wh, wl, p, bias, out, grid, st);
        case 128: return launch_tc<128, NPASS>(ah, al, wh, wl, p, bias, out, grid, st);
        case 256: return launch_tc<256, NPASS>(ah, al, wh, wl, p, bias, out, grid, st);
    }
    dsr_set_error("conv_tc: unsupported BLOCK_N %d", bn);
    return DSR_ERR_UNSUPPORTED;
}

static int tc_prep_launch(const float* x, int N, int H, int W, int C, const float* prm, const NormFin fin, int act, float slope, int pad,
                          int pad_mode, int layout, int Cp, void* A_hi, void* A_lo, void* A_bf, int Ha, int Wa, int Ca, int f16,
                          double* csum, int csum_reps, void* stream) {
    DSR_REQUIRE(!csum || (csum_reps >= 1 && csum_reps <= 64), "csum_reps: 1..64 replicas of the channel-sum row");
    DSR_REQUIRE(x && A_hi && N > 0 && H > 0 && W > 0 && C > 0, "bad arguments");
    DSR_REQUIRE(((Ca & 63) == 0 || (Ca == 8 && layout == DSR_TC_LAYOUT_NORMAL)) && (Cp & 7) == 0 && Cp >= C,
                "Ca must be a multiple of 64 (or 8 for the compact first-layer operand) and Cp a multiple of 8 >= C");
    DSR_REQUIRE(pad_mode != DSR_PAD_REFLECT || (pad < H && pad < W), "reflect padding needs pad < size");
    DSR_REQUIRE((layout == DSR_TC_LAYOUT_NORMAL && Ca >= Cp) || (layout == DSR_TC_LAYOUT_PAIR && Ca % Cp == 0 && Ca >= 2 * Cp) ||
                    (layout == DSR_TC_LAYOUT_S2D && Ca == 4 * Cp), "layout / channel mismatch");
    DSR_REQUIRE(!((uintptr_t)A_hi & 15) && !((uintptr_t)A_lo & 15), "operand buffers must be 16-byte aligned");
    DSR_REQUIRE(!csum || (layout != DSR_TC_LAYOUT_PAIR && (pad == 0 || pad_mode == DSR_PAD_ZERO)),
                "channel sums need every source element to be written exactly once (no pixel groups, zero padding)");
    DSR_REQUIRE((long)N * (H + 2 * pad) * (W + 2 * pad) < (1L << 31) && C <= 8192, "tensor too large for 32-bit pixel indices");
    const long rows = (long)N * Ha;
    // with bias-gradient sums every block ends with C fp64 atomics on the same C addresses: fewer, longer blocks
    const long cap = (long)dsr_num_sms() * ((csum && csum_reps < 4) ? 3 : (csum ? 6 : 8));
    // balanced: every block gets the same number of rows (a capped grid of 1184 blocks over 1560 rows has a two-row makespan
    // at three quarters of the blocks' worth of work)
    const long per = (rows + cap - 1) / cap;
    const int grid = (int)((rows + per - 1) / per);
    const size_t Cs = ((size_t)C + 3) & ~(size_t)3;
    const size_t smem = (((prm || fin.sums) ? 3 * Cs : 0) + (csum ? (size_t)C : 0)) * sizeof(float);
    // floor(n / d) == umulhi(n, 2^32 / d + 1) whenever n * d < 2^32
    auto magic = [](unsigned d) { return d <= 1 ? 0u : (unsigned)((1ull << 32) / d + 1); };   // 0: divisor 1
    DSR_REQUIRE((unsigned long long)Wa * (Ca >> 3) * (Ca >> 3) < (1ull << 32) && (unsigned long long)Ca * Cp < (1ull << 32) &&
                    (unsigned long long)rows * Ha < (1ull << 32), "tensor too large for the magic-number index divisions");
    PrepCat cat;
    cat.n = 0;
    for (int k = 0; k < 4; ++k) { cat.p[k] = nullptr; cat.c[k] = 0; }
    DSR_REQUIRE(!((uintptr_t)A_bf & 15), "operand buffers must be 16-byte aligned");
    {
        const int cg = Ca >> 3;
        const char* e = getenv("DSR_PREP_FAST");
        if ((!e || atoi(e) != 0) && (C & 7) == 0 && cg >= 1 && cg <= 256 && (cg & (cg - 1)) == 0 && !((uintptr_t)x & 15) &&
            (layout == DSR_TC_LAYOUT_NORMAL || Cp <= Ca)) {
            const bool has_prm = prm || fin.sums;
#define PREP_FAST(P, S) tc_prep_fast_kernel<P, S, false><<<grid, 256, smem, ST(stream)>>>(x, nullptr, nullptr, N, H, W, C, prm, act, slope, pad, pad_mode, layout, Cp, \
                (unsigned short*)A_hi, (unsigned short*)A_lo, (unsigned short*)A_bf, Ha, Wa, Ca, f16, csum, csum ? csum_reps : 1, magic((unsigned)Ha), fin)
            if (has_prm) { if (csum) PREP_FAST(true, true); else PREP_FAST(true, false); }
            else { if (csum) PREP_FAST(false, true); else PREP_FAST(false, false); }
#undef PREP_FAST
            return dsr_check_launch("tc_prep (fast)");
        }
    }
    tc_prep_kernel<<<grid, 256, smem, ST(stream)>>>(x, cat, N, H, W, C, prm, act, slope, pad, pad_mode, layout, Cp,
                                                    (unsigned short*)A_hi, (unsigned short*)A_lo, (unsigned short*)A_bf, Ha, Wa, Ca, f16, csum, csum ? csum_reps : 1,
                                                    magic((unsigned)(Ca >> 3)), magic((unsigned)Cp), magic((unsigned)Ha), fin);
    return dsr_check_launch("tc_prep");
}

extern "C" int dsr_tc_prep(const float* x, int N, int H, int W, int C, const float* prm, int act, float slope, int pad,
                           int pad_mode, int layout, int Cp, void* A_hi, void* A_lo, void* A_bf, int Ha, int Wa, int Ca, int f16,
                           double* csum, int csum_reps, void* stream) {
    NormFin fin = {};
    return tc_prep_launch(x, N, H, W, C, prm, fin, act, slope, pad, pad_mode, layout, Cp, A_hi, A_lo, A_bf, Ha, Wa, Ca, f16, csum,
                          csum_reps, stream);
}

// dsr_tc_prep with dsr_norm_finalize folded in: the operand preparation derives (mean, scale, shift) from the raw channel
// sums itself (same arithmetic, bit-identical) and writes prm_out for the readers of the backward pass - one launch less
// per normalised layer on the critical path of the step
extern "C" int dsr_tc_prep_fin(const float* x, int N, int H, int W, int C, const double* sums, int groups, const float* gamma,
                               const float* beta, float eps, float* prm_out, int act, float slope, int pad, int pad_mode,
                               int layout, int Cp, void* A_hi, void* A_lo, void* A_bf, int Ha, int Wa, int Ca, int f16,
                               double* csum, int csum_reps, void* stream) {
    DSR_REQUIRE(sums && N > 0 && C > 0 && H > 0 && W > 0 && groups >= 0 && (groups == 0 || C % groups == 0), "bad normalisation arguments");
    NormFin fin;
    fin.sums = sums; fin.gamma = gamma; fin.beta = beta; fin.prm_out = prm_out; fin.P = (long)H * W; fin.groups = groups; fin.eps = eps;
    return tc_prep_launch(x, N, H, W, C, nullptr, fin, act, slope, pad, pad_mode, layout, Cp, A_hi, A_lo, A_bf, Ha, Wa, Ca, f16, csum,
                          csum_reps, stream);
}

// Normalisation layer of a residual block + operand of the NEXT convolution in one pass:
//   y = act(norm(x)) + res  (fp32 NHWC, what dsr_norm_apply_fwd_fin writes; res optional)   and   A = arranged 16-bit planes of y
// (what dsr_tc_prep makes of y afterwards).  Statistics are finalised in the same launch (prm_out as dsr_norm_finalize).
// Covers the shapes of the fast kernel only (C % 8 == 0, Ca / 8 a power of two <= 256, NORMAL or S2D layout); anything else
// returns DSR_ERR_UNSUPPORTED and the caller runs the two passes.  models/networks.py:478-480 (ResnetBlock: out = x + conv_block(x)).
extern "C" int dsr_tc_prep_norm_res(const float* x, int N, int H, int W, int C, const double* sums, int groups, const float* gamma,
                                    const float* beta, float eps, float* prm_out, int act, const float* res, float* y_out, int pad,
                                    int pad_mode, int layout, int Cp, void* A_hi, void* A_lo, void* A_bf, int Ha, int Wa, int Ca,
                                    int f16, void* stream) {
    DSR_REQUIRE(x && sums && y_out && A_hi && N > 0 && H > 0 && W > 0 && C > 0, "bad arguments");
    DSR_REQUIRE(groups >= 0 && (groups == 0 || C % groups == 0), "C must be a multiple of groups");
    DSR_REQUIRE((Ca & 63) == 0 && (Cp & 7) == 0 && Cp >= C, "Ca must be a multiple of 64 and Cp a multiple of 8 >= C");
    DSR_REQUIRE(pad_mode != DSR_PAD_REFLECT || (pad < H && pad < W), "reflect padding needs pad < size");
    DSR_REQUIRE((layout == DSR_TC_LAYOUT_NORMAL && Ca >= Cp) || (layout == DSR_TC_LAYOUT_S2D && Ca == 4 * Cp), "layout / channel mismatch");
    DSR_REQUIRE(!((uintptr_t)A_hi & 15) && !((uintptr_t)A_lo & 15) && !((uintptr_t)A_bf & 15), "operand buffers must be 16-byte aligned");
    DSR_REQUIRE((long)N * (H + 2 * pad) * (W + 2 * pad) < (1L << 31) && C <= 8192, "tensor too large for 32-bit pixel indices");
    const int cg = Ca >> 3;
    if ((C & 7) || cg > 256 || (cg & (cg - 1)) || ((uintptr_t)x & 15) || ((uintptr_t)res & 15) || ((uintptr_t)y_out & 15)) {
        dsr_set_error("dsr_tc_prep_norm_res: shape not covered by the fused pass");
        return DSR_ERR_UNSUPPORTED;
    }
    // every source pixel must own an arranged position: the arranged grid has to cover the whole padded image
    DSR_REQUIRE(layout == DSR_TC_LAYOUT_NORMAL ? (Ha >= H + 2 * pad && Wa >= W + 2 * pad) : (2 * Ha >= H + 2 * pad && 2 * Wa >= W + 2 * pad),
                "arranged grid smaller than the padded image");
    NormFin fin;
    fin.sums = sums; fin.gamma = gamma; fin.beta = beta; fin.prm_out = prm_out; fin.P = (long)H * W; fin.groups = groups; fin.eps = eps;
    const long rows = (long)N * Ha;
    const long cap = (long)dsr_num_sms() * 8;
    const long per = (rows + cap - 1) / cap;
    const int grid = (int)((rows + per - 1) / per);
    const size_t Cs = ((size_t)C + 3) & ~(size_t)3;
    DSR_REQUIRE((unsigned long long)rows * Ha < (1ull << 32), "tensor too large for the magic-number index divisions");
    auto magic = [](unsigned d) { return d <= 1 ? 0u : (unsigned)((1ull << 32) / d + 1); };
    tc_prep_fast_kernel<true, false, true><<<grid, 256, 3 * Cs * sizeof(float), ST(stream)>>>(
        x, res, y_out, N, H, W, C, nullptr, act, 0.f, pad, pad_mode, layout, Cp, (unsigned short*)A_hi, (unsigned short*)A_lo,
        (unsigned short*)A_bf, Ha, Wa, Ca, f16, nullptr, 1, magic((unsigned)Ha), fin);
    return dsr_check_launch("tc_prep_norm_res");
}

// InstanceNorm backward apply + the arranged dY operand of the convolution in front of the norm layer (see tc_prep_inbwd_kernel).
// x / dy / dx: (N, H, W, C) fp32 NHWC; prm = (mean, rstd, .) per (n, c); sums2 from dsr_in_bwd_sums.  Shapes outside the fast
// scheme return DSR_ERR_UNSUPPORTED (the caller runs dsr_in_bwd_apply, the consumer its own dsr_tc_prep).
// models/networks.py:30, :380-381, :480 (the InstanceNorm2d layers), :378-379, :413-414 (the convolutions in front of them).
extern "C" int dsr_tc_prep_in_bwd(const float* x, const float* dy, const float* prm, const double* sums2, float* dx, int N, int H,
                                  int W, int C, int act, int pad, int layout, int Cp, void* A_hi, void* A_lo, int Ha, int Wa, int Ca,
                                  int f16, double* csum, int csum_reps, void* stream) {
    DSR_REQUIRE(x && dy && prm && sums2 && dx && A_hi && N > 0 && H > 0 && W > 0 && C > 0 && pad >= 0, "bad arguments");
    DSR_REQUIRE(act == DSR_ACT_NONE || act == DSR_ACT_RELU, "InstanceNorm backward fuses ReLU only");
    DSR_REQUIRE(!csum || (csum_reps >= 1 && csum_reps <= 64), "csum_reps: 1..64 replicas of the channel-sum row");
    DSR_REQUIRE((Ca & 63) == 0 && (Cp & 7) == 0 && Cp >= C, "Ca must be a multiple of 64 and Cp a multiple of 8 >= C");
    DSR_REQUIRE((layout == DSR_TC_LAYOUT_NORMAL && Ca >= Cp) || (layout == DSR_TC_LAYOUT_S2D && Ca == 4 * Cp), "layout / channel mismatch");
    DSR_REQUIRE(!((uintptr_t)A_hi & 15) && !((uintptr_t)A_lo & 15), "operand buffers must be 16-byte aligned");
    DSR_REQUIRE((long)N * (H + 2 * pad) * (W + 2 * pad) < (1L << 31) && C <= 8192, "tensor too large for 32-bit pixel indices");
    const int cg = Ca >> 3;
    if ((C & 7) || cg > 256 || (cg & (cg - 1)) || ((uintptr_t)x & 15) || ((uintptr_t)dy & 15) || ((uintptr_t)dx & 15)) {
        dsr_set_error("dsr_tc_prep_in_bwd: shape not covered by the fused pass");
        return DSR_ERR_UNSUPPORTED;
    }
    // every source pixel must own an arranged position: dx is written from there
    DSR_REQUIRE(layout == DSR_TC_LAYOUT_NORMAL ? (Ha >= H + 2 * pad && Wa >= W + 2 * pad) : (2 * Ha >= H + 2 * pad && 2 * Wa >= W + 2 * pad),
                "arranged grid smaller than the padded image");
    const long rows = (long)N * Ha;
    const long cap = (long)dsr_num_sms() * ((csum && csum_reps < 4) ? 3 : 6);
    const long per = (rows + cap - 1) / cap;
    const int grid = (int)((rows + per - 1) / per);
    const size_t Cs = ((size_t)C + 3) & ~(size_t)3;
    const size_t smem = (4 * Cs + (csum ? (size_t)C : 0)) * sizeof(float);
    if (smem > 48 * 1024) {
        dsr_set_error("dsr_tc_prep_in_bwd: %d channels need more than 48 KB of per-sample constants", C);
        return DSR_ERR_UNSUPPORTED;
    }
    DSR_REQUIRE((unsigned long long)rows * Ha < (1ull << 32), "tensor too large for the magic-number index divisions");
    auto magic = [](unsigned d) { return d <= 1 ? 0u : (unsigned)((1ull << 32) / d + 1); };
    if (csum)
        tc_prep_inbwd_kernel<true><<<grid, 256, smem, ST(stream)>>>(dy, x, dx, N, H, W, C, prm, sums2, act, pad, layout, Cp,
            (unsigned short*)A_hi, (unsigned short*)A_lo, Ha, Wa, Ca, f16, csum, csum_reps, magic((unsigned)Ha));
    else
        tc_prep_inbwd_kernel<false><<<grid, 256, smem, ST(stream)>>>(dy, x, dx, N, H, W, C, prm, sums2, act, pad, layout, Cp,
            (unsigned short*)A_hi, (unsigned short*)A_lo, Ha, Wa, Ca, f16, nullptr, 1, magic((unsigned)Ha));
    return dsr_check_launch("tc_prep_in_bwd");
}

// the same preparation over torch.cat((x0, x1, x2, x3), dim=1) without materialising the concatenation
// (models/main_model.py:305-306: the 261-channel Task input; models/networks.py:629: the U-Net skip connections)
extern "C" int dsr_tc_prep_cat(const float* x0, int C0, const float* x1, int C1, const float* x2, int C2, const float* x3, int C3,
                               int N, int H, int W, int pad, int pad_mode, int layout, int Cp, void* A_hi, void* A_lo, void* A_bf,
                               int Ha, int Wa, int Ca, int f16, void* stream) {
    PrepCat cat;
    const float* ps[4] = {x0, x1, x2, x3};
    const int cs[4] = {C0, C1, C2, C3};
    cat.n = 0;
    int C = 0;
    for (int k = 0; k < 4; ++k) {
        cat.p[k] = nullptr; cat.c[k] = 0;
        if (ps[k] && cs[k] > 0) {
            DSR_REQUIRE(cat.n == k, "sources must be given in order without gaps");
            cat.p[k] = ps[k]; cat.c[k] = cs[k]; C += cs[k]; cat.n = k + 1;
        }
    }
    DSR_REQUIRE(cat.n >= 1 && A_hi && N > 0 && H > 0 && W > 0, "bad arguments");
    DSR_REQUIRE(C <= 8 * PREP_CAT_MAX_GROUPS, "too many concatenated channels");
    DSR_REQUIRE((Ca & 63) == 0 && (Cp & 7) == 0 && Cp >= C, "Ca must be a multiple of 64 and Cp a multiple of 8 >= the channel total");
    DSR_REQUIRE(pad_mode != DSR_PAD_REFLECT || (pad < H && pad < W), "reflect padding needs pad < size");
    DSR_REQUIRE((layout == DSR_TC_LAYOUT_NORMAL && Ca >= Cp) || (layout == DSR_TC_LAYOUT_PAIR && Ca % Cp == 0 && Ca >= 2 * Cp) ||
                    (layout == DSR_TC_LAYOUT_S2D && Ca == 4 * Cp), "layout / channel mismatch");
    DSR_REQUIRE(!((uintptr_t)A_hi & 15) && !((uintptr_t)A_lo & 15), "operand buffers must be 16-byte aligned");
    DSR_REQUIRE((long)N * (H + 2 * pad) * (W + 2 * pad) < (1L << 31) && C <= 8192, "tensor too large for 32-bit pixel indices");
    const long rows = (long)N * Ha;
    const long cap = (long)dsr_num_sms() * 8;
    const int grid = (int)(rows < cap ? rows : cap);
    auto magic = [](unsigned d) { return d <= 1 ? 0u : (unsigned)((1ull << 32) / d + 1); };
    DSR_REQUIRE((unsigned long long)Wa * (Ca >> 3) * (Ca >> 3) < (1ull << 32) && (unsigned long long)Ca * Cp < (1ull << 32) &&
                    (unsigned long long)rows * Ha < (1ull << 32), "tensor too large for the magic-number index divisions");
    DSR_REQUIRE(!((uintptr_t)A_bf & 15), "operand buffers must be 16-byte aligned");
    tc_prep_kernel<<<grid, 256, 0, ST(stream)>>>(x0, cat, N, H, W, C, nullptr, DSR_ACT_NONE, 0.f, pad, pad_mode, layout, Cp,
                                                 (unsigned short*)A_hi, (unsigned short*)A_lo, (unsigned short*)A_bf, Ha, Wa, Ca, f16, nullptr, 1,
                                                 magic((unsigned)(Ca >> 3)), magic((unsigned)Cp), magic((unsigned)Ha), NormFin{});
    return dsr_check_launch("tc_prep_cat");
}

extern "C" int dsr_tc_pack_weight(const float* w, int D0, int D1, int R, int S, int variant, int Cp, int phase_a, int phase_b,
                                  int pad, int Cout, int T, int Ca, void* W_hi, void* W_lo, int f16, float wscale,
                                  void* stream) {
    DSR_REQUIRE(w && W_hi && T > 0 && (Ca & 63) == 0, "bad arguments");
    DSR_REQUIRE(!((uintptr_t)W_hi & 15) && !((uintptr_t)W_lo & 15), "packed weight buffers must be 16-byte aligned");
    long total = (long)Cout * T * (Ca / 8);
    DSR_REQUIRE(phase_a >= 0 || variant == DSR_TC_W_CONVT_PH, "phase -1 (all four phases, stacked rows) is for the transposed-conv variant");
    {
        // big layers: shared-memory tiled form (DSR_PACK_TILED=0 keeps the gather kernel, for A/B tests)
        const char* e = getenv("DSR_PACK_TILED");
        const bool convT = variant == DSR_TC_W_CONVT_PH || variant == DSR_TC_W_CONV_DGRAD;
        const bool kind_ok = variant == DSR_TC_W_CONV || variant == DSR_TC_W_CONV_S2D || convT;
        const int RS = R * S;
        const int Texp = (variant == DSR_TC_W_CONV_S2D || variant == DSR_TC_W_CONVT_PH) ? 4 : RS;
        if ((!e || atoi(e) != 0) && kind_ok && R == S && (RS == 9 || RS == 16) && T == Texp && (long)D0 * D1 * RS >= (1L << 16) &&
            (Cp & 7) == 0) {
#define PACK_TILED(V, K) pack_tiled_launch<V, K>(w, D0, D1, Cp, phase_a, phase_b, pad, Cout, Ca, W_hi, W_lo, f16, wscale, ST(stream))
            if (variant == DSR_TC_W_CONV) { if (RS == 9) PACK_TILED(DSR_TC_W_CONV, 9); else PACK_TILED(DSR_TC_W_CONV, 16); }
            else if (variant == DSR_TC_W_CONV_S2D) { if (RS == 9) PACK_TILED(DSR_TC_W_CONV_S2D, 9); else PACK_TILED(DSR_TC_W_CONV_S2D, 16); }
            else if (variant == DSR_TC_W_CONV_DGRAD) { if (RS == 9) PACK_TILED(DSR_TC_W_CONV_DGRAD, 9); else PACK_TILED(DSR_TC_W_CONV_DGRAD, 16); }
            else { if (RS == 9) PACK_TILED(DSR_TC_W_CONVT_PH, 9); else PACK_TILED(DSR_TC_W_CONVT_PH, 16); }
#undef PACK_TILED
            return dsr_check_launch("tc_pack_weight (tiled)");
        }
    }
    tc_pack_weight_kernel<<<dim3(dsr_grid(total, 256, phase_a < 0 ? 2 : 8), phase_a < 0 ? 4 : 1), 256, 0, ST(stream)>>>(w, D0, D1, R, S, variant, Cp, phase_a, phase_b, pad, Cout, T,
                                                                       Ca, (unsigned short*)W_hi, (unsigned short*)W_lo, f16, wscale);
    return dsr_check_launch("tc_pack_weight");
}

extern "C" int dsr_tc_gemm(const void* A_hi, const void* A_lo, int N, int Ha, int Wa, int Ca, const void* W_hi, const void* W_lo,
                           int Cout, int T, const int* tap_dr, const int* tap_ds, int a_off_h, int a_off_w, int Ht, int Wt,
                           const float* bias, float* out, int Ho, int Wo, int os, int ph, int pw, int nphase, int act, int npass,
                           int split_k, int f16, float out_scale, void* stream) {
    DSR_REQUIRE(A_hi && W_hi && out && tap_dr && tap_ds, "null pointer");
    DSR_REQUIRE(nphase == 1 || (nphase == 4 && os == 2 && ph == 0 && pw == 0), "phases: 1, or 4 with output stride 2");
    DSR_REQUIRE(npass >= 1 && npass <= 3 && (npass < 2 || A_lo) && (npass < 3 || W_lo), "bad precision mode");
    DSR_REQUIRE(T >= 1 && T <= TC_MAX_TAPS && (Ca & 63) == 0 && Cout >= 1, "bad GEMM shape");
    DSR_REQUIRE(!((uintptr_t)A_hi & 15) && !((uintptr_t)W_hi & 15) && !((uintptr_t)out & 15), "buffers must be 16-byte aligned");
    TcParams p;
    p.N = N; p.Ht = Ht; p.Wt = Wt; p.Ca = Ca; p.T = T; p.ah = a_off_h; p.aw = a_off_w;
    p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.os = os; p.ph = ph; p.pw = pw; p.act = act;
    p.f16 = f16; p.out_scale = out_scale;
    for (int t = 0; t < T; ++t) { p.dr[t] = (signed char)tap_dr[t]; p.ds[t] = (signed char)tap_ds[t]; }
    // tile shape: TW x TH x TN = 128 rows
    int TW = Wt >= 16 ? 16 : pow2_ceil(Wt);
    int TH = 128 / TW;
    if (TH > pow2_ceil(Ht)) TH = pow2_ceil(Ht);
    int TN = 128 / (TW * TH);
    p.TW = TW; p.TH = TH; p.TN = TN;
    p.tiles_w = dsr_cdiv(Wt, TW); p.tiles_h = dsr_cdiv(Ht, TH);
    int tiles_n = dsr_cdiv(N, TN);
    int bn = Cout >= 256 ? 256 : (Cout > 64 ? 128 : (Cout > 32 ? 64 : (Cout > 16 ? 32 : 16)));
    if (npass == 3 && bn == 256) bn = 128;          // keep >= 2 pipeline stages in 200 KB of smem
    int tiles_co = dsr_cdiv(Cout, bn);
    p.ksteps = T * (Ca / 64);
    long ctas = (long)p.tiles_w * p.tiles_h * tiles_n * tiles_co * nphase;
    int splits = 1;
    if (split_k < 0) {              // auto: only when the grid would leave most SMs idle
        if (ctas * 2 <= dsr_num_sms() && p.ksteps >= 16) {
            splits = (int)(dsr_num_sms() / ctas);
            if (splits > p.ksteps / 4) splits = p.ksteps / 4;
            if (splits < 1) splits = 1;
        }
    } else if (split_k > 1) splits = split_k;
    if (splits > 1 && act != DSR_ACT_NONE) splits = 1;

    p.ksteps_per_split = dsr_cdiv(p.ksteps, splits);
    splits = dsr_cdiv(p.ksteps, p.ksteps_per_split);
    if (splits > 1) {
        if (cudaMemsetAsync(out, 0, (size_t)N * Ho * Wo * Cout * sizeof(float), ST(stream)) != cudaSuccess) {
            dsr_set_error("conv_tc: memset failed"); return DSR_ERR_CUDA;
        }
        DSR_REQUIRE(os == 1 || nphase == 4, "split-K of a single output phase would zero the other phases");
    }
    CUtensorMap mah, mal, mwh, mwl;
    cuuint64_t adims[4] = {(cuuint64_t)Ca, (cuuint64_t)Wa, (cuuint64_t)Ha, (cuuint64_t)N};
    cuuint64_t astr[3] = {(cuuint64_t)Ca * 2, (cuuint64_t)Wa * Ca * 2, (cuuint64_t)Ha * Wa * Ca * 2};
    cuuint32_t abox[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
    int rc = encode_map(&mah, A_hi, 4, adims, astr, abox);
    if (rc) return rc;
    mal = mah;
    if (npass >= 2 && (rc = encode_map(&mal, A_lo, 4, adims, astr, abox))) return rc;
    cuuint64_t wdims[2] = {(cuuint64_t)T * Ca, (cuuint64_t)Cout * nphase};     // phase weight matrices stacked along rows
    cuuint64_t wstr[1] = {(cuuint64_t)T * Ca * 2};
    cuuint32_t wbox[2] = {64, (cuuint32_t)bn};
    if ((rc = encode_map(&mwh, W_hi, 2, wdims, wstr, wbox))) return rc;
    mwl = mwh;
    if (npass >= 3 && (rc = encode_map(&mwl, W_lo, 2, wdims, wstr, wbox))) return rc;
    p.nphase = nphase; p.tiles_per_phase = p.tiles_w * p.tiles_h * tiles_n;
    dim3 grid((unsigned)(p.tiles_per_phase * nphase), (unsigned)tiles_co, (unsigned)splits);
    if (npass == 1) return dispatch_n<1>(bn, mah, mal, mwh, mwl, p, bias, out, grid, ST(stream));
    if (npass == 2) return dispatch_n<2>(bn, mah, mal, mwh, mwl, p, bias, out, grid, ST(stream));
    return dispatch_n<3>(bn, mah, mal, mwh, mwl, p, bias, out, grid, ST(stream));
}

template <int BLOCK_N, int NPASS>
static int launch_wg(const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& ah, const CUtensorMap& al,
                     const WgParams& p, float* dWp, dim3 grid, cudaStream_t st) {
    using Cfg = WgCfg<BLOCK_N, NPASS>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel<BLOCK_N, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) {
            dsr_set_error("wgrad_tc: cannot raise dynamic shared memory to %d", Cfg::SMEM);
            return DSR_ERR_CUDA;
        }
        attr = true;
    }
    wgrad_tc_kernel<BLOCK_N, NPASS><<<grid, 192, Cfg::SMEM, st>>>(mh, ml, ah, al, p, dWp);
    return dsr_check_launch("wgrad_tc");
}
template <int NPASS>
static int dispatch_wg(int bn, const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& ah, const CUtensorMap& al,
                       const WgParams& p, float* dWp, dim3 grid, cudaStream_t st) {
    switch (bn) {
        case 64: return launch_wg<64, NPASS>(mh, ml, ah, al, p, dWp, grid, st);
        case 128: return launch_wg<128, NPASS>(mh, ml, ah, al, p, dWp, grid, st);
        case 256: return launch_wg<256, NPASS>(mh, ml, ah, al, p, dWp, grid, st);
    }
    dsr_set_error("wgrad_tc: unsupported BLOCK_N %d", bn);
    return DSR_ERR_UNSUPPORTED;
}

extern "C" int dsr_tc_wgrad(const void* M_hi, const void* M_lo, int N, int Hm, int Wm, int Cm, int Cm_real, int m_off_h,
                            int m_off_w, const void* A_hi, const void* A_lo, int Ha, int Wa, int Ca, int T,
                            const int* tap_dr, const int* tap_ds, int a_off_h, int a_off_w, int Hb, int Wb, float* dWp,
                            int npass, int f16, float out_scale, int split_k, void* stream) {
    DSR_REQUIRE(M_hi && A_hi && dWp && tap_dr && tap_ds, "null pointer");
    DSR_REQUIRE(npass >= 1 && npass <= 3 && (npass < 2 || M_lo) && (npass < 3 || A_lo), "bad precision mode");
    DSR_REQUIRE(T >= 1 && T <= TC_MAX_TAPS && (Ca & 63) == 0 && (Cm & 63) == 0 && Cm_real >= 1 && Cm_real <= Cm, "bad GEMM shape");
    WgParams p;
    p.N = N; p.Hb = Hb; p.Wb = Wb; p.mh = m_off_h; p.mw = m_off_w; p.ah = a_off_h; p.aw = a_off_w;
    p.Cm_real = Cm_real; p.Ca = Ca; p.T = T; p.f16 = f16; p.out_scale = out_scale;
    for (int t = 0; t < T; ++t) { p.dr[t] = (signed char)tap_dr[t]; p.ds[t] = (signed char)tap_ds[t]; }
    int TW = Wb >= 16 ? 16 : pow2_ceil(Wb);
    int TH = 64 / TW;
    if (TH > pow2_ceil(Hb)) TH = pow2_ceil(Hb);
    int TN = 64 / (TW * TH);
    p.TW = TW; p.TH = TH; p.TN = TN;
    p.tiles_w = dsr_cdiv(Wb, TW); p.tiles_h = dsr_cdiv(Hb, TH);
    p.tiles_total = p.tiles_w * p.tiles_h * dsr_cdiv(N, TN);
    int bn = Ca >= 256 ? 256 : (Ca >= 128 ? 128 : 64);
    if (npass == 3 && bn == 256) bn = 128;
    p.n_tiles_c = dsr_cdiv(Ca, bn);
    int tiles_m = dsr_cdiv(Cm_real, 128);
    long ctas = (long)T * tiles_m * p.n_tiles_c;
    int splits = 1;
    if (split_k < 0) {
        splits = (int)((2L * dsr_num_sms() + ctas - 1) / ctas);
        if (splits > p.tiles_total / 4) splits = p.tiles_total / 4;
        if (splits < 1) splits = 1;
    } else if (split_k > 1) splits = split_k;
    p.tiles_per_split = dsr_cdiv(p.tiles_total, splits);
    splits = dsr_cdiv(p.tiles_total, p.tiles_per_split);
    if (splits > 1 && cudaMemsetAsync(dWp, 0, (size_t)Cm_real * T * Ca * sizeof(float), ST(stream)) != cudaSuccess) {
        dsr_set_error("wgrad_tc: memset failed"); return DSR_ERR_CUDA;
    }
    CUtensorMap mmh, mml, mah, mal;
    cuuint64_t mdims[4] = {(cuuint64_t)Cm, (cuuint64_t)Wm, (cuuint64_t)Hm, (cuuint64_t)N};
    cuuint64_t mstr[3] = {(cuuint64_t)Cm * 2, (cuuint64_t)Wm * Cm * 2, (cuuint64_t)Hm * Wm * Cm * 2};
    cuuint64_t adims[4] = {(cuuint64_t)Ca, (cuuint64_t)Wa, (cuuint64_t)Ha, (cuuint64_t)N};
    cuuint64_t astr[3] = {(cuuint64_t)Ca * 2, (cuuint64_t)Wa * Ca * 2, (cuuint64_t)Ha * Wa * Ca * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
    int rc = encode_map(&mmh, M_hi, 4, mdims, mstr, box);
    if (rc) return rc;
    mml = mmh;
    if (npass >= 2 && (rc = encode_map(&mml, M_lo, 4, mdims, mstr, box))) return rc;
    if ((rc = encode_map(&mah, A_hi, 4, adims, astr, box))) return rc;
    mal = mah;
    if (npass >= 3 && (rc = encode_map(&mal, A_lo, 4, adims, astr, box))) return rc;
    dim3 grid((unsigned)T, (unsigned)(tiles_m * p.n_tiles_c), (unsigned)splits);
    if (npass == 1) return dispatch_wg<1>(bn, mmh, mml, mah, mal, p, dWp, grid, ST(stream));
    if (npass == 2) return dispatch_wg<2>(bn, mmh, mml, mah, mal, p, dWp, grid, ST(stream));
    return dispatch_wg<3>(bn, mmh, mml, mah, mal, p, dWp, grid, ST(stream));
}

// Tiled inverse of the packing for the big layers (variants CONV and CONV_S2D, 3x3 / 4x4 kernels): a CTA reads the
// [t][c0 .. c0 + 128) runs of one packed row contiguously, transposes them through shared memory and writes (or accumulates
// into) the parameter's own contiguous [c][r][s] run.  The gather kernel above reads 4-byte elements Ca floats apart.
template <int VARIANT, int RS>
__global__ void __launch_bounds__(256)
tc_unpack_wgrad_tiled_kernel(const float* __restrict__ dWp, int D0, int D1, int Cp, int Ca, float* __restrict__ grad, int accumulate,
                             int nsplit, long split_stride) {
    constexpr int S = RS == 9 ? 3 : 4, R = S, TC = 128, RSP = RS + 1;
    constexpr int NAB = VARIANT == DSR_TC_W_CONV_S2D ? 4 : 1, T = VARIANT == DSR_TC_W_CONV_S2D ? 4 : RS;
    __shared__ float gt[TC * RSP];
    const int nc_tiles = (D1 + TC - 1) / TC, tid = threadIdx.x;
    for (int tile = blockIdx.x; tile < D0 * nc_tiles; tile += gridDim.x) {
        const int d0 = tile / nc_tiles, c0 = (tile - d0 * nc_tiles) * TC, nc = min(TC, D1 - c0);
        const float* src = dWp + (long)d0 * ((long)T * Ca);
        __syncthreads();
        for (int i = tid; i < T * NAB * TC; i += 256) {
            const int c_l = i % TC, u = i / TC, ab = u % NAB, t = u / NAB;
            int r, sx;
            if (NAB == 4) { r = 2 * (t >> 1) + (ab >> 1); sx = 2 * (t & 1) + (ab & 1); } else { r = t / S; sx = t - r * S; }
            if (c_l < nc && r < R && sx < S) {
                const float* e = src + (long)t * Ca + ab * Cp + c0 + c_l;
                float v = __ldg(e);
                for (int z = 1; z < nsplit; ++z) v += __ldg(e + z * split_stride);     // K-split slabs, fixed order
                gt[c_l * RSP + r * S + sx] = v;
            }
        }
        __syncthreads();
        float* dst = grad + ((long)d0 * D1 + c0) * RS;
        for (int i = tid; i < nc * RS; i += 256) {
            const int c_l = i / RS;
            const float v = gt[c_l * RSP + (i - c_l * RS)];
            if (accumulate) dst[i] += v; else dst[i] = v;
        }
    }
}
// nsplit slabs of [D0][T*Ca] (dsr_tc_wgrad2p with partial = 1) summed in slab order on the way in; nsplit = 1: one matrix
extern "C" int dsr_tc_unpack_wgrad_splits(const float* dWp, int nsplit, int D0, int D1, int R, int S, int variant, int Cp, int T,
                                          int Ca, float* grad, int accumulate, void* stream) {
    DSR_REQUIRE(dWp && grad && nsplit >= 1, "null pointer / bad split count");
    const long split_stride = (long)D0 * T * Ca;
    DSR_REQUIRE(variant == DSR_TC_W_CONV || variant == DSR_TC_W_CONV_PAIR || variant == DSR_TC_W_CONV_S2D, "unsupported variant");
    {
        const char* e = getenv("DSR_PACK_TILED");
        const int RS = R * S, Texp = variant == DSR_TC_W_CONV_S2D ? 4 : RS;
        if ((!e || atoi(e) != 0) && variant != DSR_TC_W_CONV_PAIR && R == S && (RS == 9 || RS == 16) && T == Texp &&
            (long)D0 * D1 * RS >= (1L << 16)) {
            const long tiles = (long)D0 * dsr_cdiv(D1, 128), cap = (long)dsr_num_sms() * 8;
            const int grid = (int)(tiles < cap ? tiles : cap);
#define UNPACK_TILED(V, K) tc_unpack_wgrad_tiled_kernel<V, K><<<grid, 256, 0, ST(stream)>>>(dWp, D0, D1, Cp, Ca, grad, accumulate, nsplit, split_stride)
            if (variant == DSR_TC_W_CONV) { if (RS == 9) UNPACK_TILED(DSR_TC_W_CONV, 9); else UNPACK_TILED(DSR_TC_W_CONV, 16); }
            else { if (RS == 9) UNPACK_TILED(DSR_TC_W_CONV_S2D, 9); else UNPACK_TILED(DSR_TC_W_CONV_S2D, 16); }
#undef UNPACK_TILED
            return dsr_check_launch("tc_unpack_wgrad (tiled)");
        }
    }
    tc_unpack_wgrad_kernel<<<dsr_grid((long)D0 * D1 * R * S, 256), 256, 0, ST(stream)>>>(dWp, D0, D1, R, S, variant, Cp, T, Ca, grad,
                                                                                        accumulate, nsplit, split_stride);
    return dsr_check_launch("tc_unpack_wgrad");
}
extern "C" int dsr_tc_unpack_wgrad(const float* dWp, int D0, int D1, int R, int S, int variant, int Cp, int T, int Ca,
                                   float* grad, int accumulate, void* stream) {
    return dsr_tc_unpack_wgrad_splits(dWp, 1, D0, D1, R, S, variant, Cp, T, Ca, grad, accumulate, stream);
}
