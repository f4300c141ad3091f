// Halo-resident, persistent tcgen05 implicit GEMM (second-generation convolution kernel, sm_100a).
//
// Same GEMM form as conv_tc.cu (out[n,h*os+ph,w*os+pw,co] = bias[co] + sum_t sum_c A[n,h+ah+dr_t,w+aw+ds_t,c] W[co][t*Ca+c]),
// different data movement.  Profiling the first kernel (profiles/r1_*) showed the 128x128 tile is bound by the
// L2->SM operand feed: every tap re-loads a full 128-pixel A tile.  Here
//   * an output tile is 8 (w) x 16 (h) pixels; for each 64-channel block ONE TMA box brings the whole input patch
//     (16 pixels wide x (16 + max_dr) rows, SWIZZLE_128B, 2 KB per patch row) into shared memory and every tap
//     of the window is issued straight from it: the A descriptor of tap (dr, ds) starts at patch + (dr*16+ds)*128 B
//     with a stride of 2048 B between 8-row groups (one group = the 8 pixels of an output row); an A byte now
//     crosses L2->SM once per 64-channel block instead of once per tap;
//   * CTAs are persistent (grid = min(tiles, SMs)); the accumulator is double-buffered in TMEM
//     (2 x BLOCK_N columns) so the epilogue of tile i overlaps the main loop of tile i+1;
//   * the epilogue can emit the InstanceNorm / GroupNorm statistics of its tile (per-(n, c) sum and sum of squares,
//     fp64 red.add) so the separate channel_sums pass over the activation disappears.
// Warp roles (224 threads): 0 = patch TMA producer, 1 = weight TMA producer, 2 = TMEM allocator + MMA issuer,
// 3..6 = epilogue (TMEM lane quadrant = warp % 4).
#include "tc_common.cuh"

#define TC2_MAX_TAPS 64
#define TC2_TH 16
#define TC2_TW 8

struct Tc2Params {
    int N, Ht, Wt;
    int tiles_w, tiles_h, tiles_co, total_tiles;
    int Ca, T, cblocks;
    int ah, aw, Hp, PW;       // patch = Hp rows x PW pixels x 128 B (PW = 8 + widest tap offset)
    int Ho, Wo, Cout, os, ph, pw;
    int a_mode;               // 1: compact first-layer operand - A is [pixels][8 channels] (16-byte pixel rows, NO swizzle); the
                              // K block of a kernel row = 8 horizontally adjacent pixels x 8 channels = 128 contiguous bytes
                              // starting at the tap's pixel, i.e. MMA row i starts 16 B after row i-1 (overlapping rows:
                              // no-swizzle K-major descriptor with LBO = 16 B, SBO = patch row pitch)
    int nphase, tiles_per_phase;  // 4 output phases of a stride-2 transposed conv in one launch (see conv_tc.cu)
    int act, f16, base_off_mode;
    int run_stats;            // 1: the epilogue keeps per-sample running statistics sums (see the epilogue); 0: atomics per tile (A/B)
    float out_scale;
    int n_pb, n_ws;
    unsigned patch_plane_bytes;    // Hp * PW * 128 rounded up to the 1024-byte swizzle repeat (smem placement)
    unsigned patch_tx_bytes;       // Hp * PW * 128 (what one TMA box delivers)
    signed char dr[TC2_MAX_TAPS], ds[TC2_MAX_TAPS];
    unsigned shift16[TC2_MAX_TAPS];   // per tap: (dr * PW + ds) * pixel bytes >> 4 (added to the patch descriptor's address field)
};

// K-major SWIZZLE_128B descriptor with an explicit 8-row-group stride and base offset (start address not aligned to
// the 1024-byte swizzle repeat: base_offset = (start >> 7) & 7)
__device__ __forceinline__ uint64_t make_sdesc_ex(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(base_off & 7) << 49) | ((uint64_t)2 << 61);
}

// optional per-CTA wait-time counters (clock64 ticks): [0] MMA waits patch, [1] MMA waits weights, [2] MMA waits a free
// accumulator, [3] MMA role total, [4] epilogue waits accumulator, [5] epilogue role total, [6] patch producer waits a
// free buffer, [7] weight producer waits a free stage.  Enabled by dsr_tc2_set_debug (profiling aid).
__device__ long long* g_tc2_dbg = nullptr;
#define TC2_TIMED_WAIT(slot, bar, parity)                 \
    do {                                                  \
        if (dbg) {                                        \
            const long long t0_ = clock64();              \
            mbar_wait(bar, parity);                       \
            dbg_acc[slot] += clock64() - t0_;             \
        } else {                                          \
            mbar_wait(bar, parity);                       \
        }                                                 \
    } while (0)

template <int BLOCK_N, int NPASS>
__global__ void __launch_bounds__(224, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
                const __grid_constant__ Tc2Params p, const float* __restrict__ bias, float* __restrict__ out,
                double* __restrict__ stats) {
    constexpr int NA = NPASS >= 2 ? 2 : 1;
    constexpr int NW = NPASS >= 3 ? 2 : 1;
    constexpr uint32_t W_TILE = BLOCK_N * 128;
    // NPASS == 4 ("wide"): the three products of the parity mode in TWO MMAs per K step.  With the A operand in shared memory
    // an M = 128 MMA costs ~64 cycles whatever N <= 128 is (4 KB of A rows at 64 B/clk: wait counters r4a - the issuer is busy
    // 62..69 cycles per MMA at N = 32 and N = 64 alike), so A_hi x [W_hi | W_lo] runs as ONE MMA of N = 2 * BLOCK_N (the W_lo
    // tile follows the W_hi tile in the stage: one K-major descriptor covers both) into 2 * BLOCK_N accumulator columns, A_lo x
    // W_hi adds into the first BLOCK_N, and the epilogue adds the two halves: 8 MMAs per stage instead of 12.
    constexpr bool WIDE = NPASS == 4;
    constexpr int ACC = WIDE ? 2 * BLOCK_N : BLOCK_N;            // accumulator columns per buffer
    constexpr uint32_t TMEM_COLS = 2 * ACC < 32 ? 32 : 2 * ACC;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t patch_set = NA * p.patch_plane_bytes;
    const uint32_t w_base = smem_base + p.n_pb * patch_set;
    const uint32_t w_stage = NW * W_TILE;
    const uint32_t bar_base = w_base + p.n_ws * w_stage;
    // barriers: pf[n_pb] pe[n_pb] wf[n_ws] we[n_ws] af[2] ae[2], then the TMEM pointer
    auto pf = [&](int i) { return bar_base + 8u * i; };
    auto pe = [&](int i) { return bar_base + 8u * (p.n_pb + i); };
    auto wf = [&](int i) { return bar_base + 8u * (2 * p.n_pb + i); };
    auto we = [&](int i) { return bar_base + 8u * (2 * p.n_pb + p.n_ws + i); };
    auto af = [&](int i) { return bar_base + 8u * (2 * p.n_pb + 2 * p.n_ws + i); };
    auto ae = [&](int i) { return bar_base + 8u * (2 * p.n_pb + 2 * p.n_ws + 2 + i); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.n_pb + 2 * p.n_ws + 4);
    const uint32_t epi_base = bar_base + 512u;          // 4 KB of per-warp bias copies + 16 KB of store staging tiles
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* dbg = g_tc2_dbg;
    long long dbg_acc[3] = {0, 0, 0};
    const long long t_start = dbg ? clock64() : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA_hi); tma_prefetch_desc(&mapW_hi);
        if (NPASS >= 2) tma_prefetch_desc(&mapA_lo);
        if (NPASS >= 3) tma_prefetch_desc(&mapW_lo);
        for (int i = 0; i < p.n_pb; ++i) { mbar_init(pf(i), 1); mbar_init(pe(i), 1); }
        for (int i = 0; i < p.n_ws; ++i) { mbar_init(wf(i), 1); mbar_init(we(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(af(i), 1); mbar_init(ae(i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0) {
        // ===== patch producer: one box per (tile, 64-channel block) =====
        {
            // ring positions and barrier parities advance incrementally in all three roles (no runtime `i % n` / `i / n`: a
            // ~30-instruction dependent chain through MUFU.RCP per stage kept the MMA issuer latency-bound - conv_tc3.cu)
            int pb = 0;
            uint32_t pph = 1u;
            const bool leader = elect_one();           // ONE thread runs the whole role (see the MMA issuer below)
            if (leader)
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int phase = tile / p.tiles_per_phase;
                int r = (tile - phase * p.tiles_per_phase) / p.tiles_co;
                const int tw_i = r % p.tiles_w; r /= p.tiles_w;
                const int th_i = r % p.tiles_h;
                const int n = r / p.tiles_h;
                const int hc = th_i * TC2_TH + p.ah + (phase >> 1), wc = tw_i * TC2_TW + p.aw + (phase & 1);
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    TC2_TIMED_WAIT(0, pe(pb), pph);
                    {
                        const uint32_t dst = smem_base + pb * patch_set;
                        mbar_expect_tx(pf(pb), NA * p.patch_tx_bytes);
                        tma_load_4d(dst, &mapA_hi, pf(pb), cb << 6, wc, hc, n);
                        if (NPASS >= 2) tma_load_4d(dst + p.patch_plane_bytes, &mapA_lo, pf(pb), cb << 6, wc, hc, n);
                    }
                    if (++pb == p.n_pb) { pb = 0; pph ^= 1u; }
                }
            }
            if (dbg && leader) dbg[blockIdx.x * 16 + 6] = dbg_acc[0];
        }
    } else if (warp == 1) {
        // ===== weight producer: one [BLOCK_N x 64] tile (hi, lo) per (tile, block, tap) =====
        {
            int ws = 0;
            uint32_t wph = 1u, dst = w_base, full = wf(0), empty = we(0);
            const bool leader = elect_one();
            if (leader)
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int phase = tile / p.tiles_per_phase;
                const int co0 = ((tile - phase * p.tiles_per_phase) % p.tiles_co) * BLOCK_N + phase * p.Cout;
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    int kw = cb << 6;
                    for (int t = 0; t < p.T; ++t, kw += p.Ca) {
                        TC2_TIMED_WAIT(0, empty, wph);
                        {
                            mbar_expect_tx(full, w_stage);
                            tma_load_2d(dst, &mapW_hi, full, kw, co0);
                            if (NPASS >= 3) tma_load_2d(dst + W_TILE, &mapW_lo, full, kw, co0);
                        }
                        dst += w_stage; full += 8u; empty += 8u;
                        if (++ws == p.n_ws) { ws = 0; wph ^= 1u; dst = w_base; full = wf(0); empty = we(0); }
                    }
                }
            }
            if (dbg && leader) dbg[blockIdx.x * 16 + 7] = dbg_acc[0];
        }
    } else if (warp == 2) {
        // ===== MMA issuer: the whole warp runs the (warp-uniform) loop and the waits; one elected lane issues =====
        {
            const uint32_t idesc = make_idesc(128, BLOCK_N < 16 ? 16 : BLOCK_N, p.f16 ? 0u : 1u);
            const uint32_t idesc_w = make_idesc(128, 2 * BLOCK_N, p.f16 ? 0u : 1u);
            // A descriptor template: SWIZZLE_128B rows of 128 B (stride between the 8-pixel row groups = patch pitch), or
            // for the compact first-layer operand no swizzle, 16-byte rows, LBO = 16 B, SBO = patch pitch
            const uint64_t pdesc0 = p.a_mode
                ? ((uint64_t)1 << 16) | ((uint64_t)(((uint32_t)p.PW * 16u) >> 4) << 32) | ((uint64_t)1 << 46)
                : make_sdesc_ex(0, (uint32_t)p.PW * 128u, 0);
            const uint64_t wdesc0 = make_sdesc(0);
            int pb = 0, ws = 0, ti = 0;
            uint32_t pph = 0u, wph = 0u, full = wf(0), empty = we(0);
            const uint64_t wd_first = wdesc0 + (uint64_t)((w_base & 0x3FFFF) >> 4);     // descriptor of weight stage 0
            uint64_t wd_hi = wd_first;
            // ONE elected thread runs the whole loop, waits included: inside a region the compiler knows to be single-threaded
            // the loop state can live in uniform registers next to the UTCHMMA operands (a per-stage `if (elect_one())` in a
            // warp-wide loop cost ~12 R2UR moves, an ELECT and a BSSY / BSYNC pair per stage)
            const bool leader = elect_one();
            if (leader)
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
                const int ab = ti & 1;
                TC2_TIMED_WAIT(2, ae(ab), ((ti >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(ab * ACC);
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    TC2_TIMED_WAIT(0, pf(pb), pph);
                    const uint32_t patch_hi = smem_base + pb * patch_set;
                    const uint64_t pd_base = pdesc0 + (uint64_t)((patch_hi & 0x3FFFF) >> 4);
                    const uint32_t pe_bar = pe(pb);
                    for (int t = 0; t < p.T; ++t) {
                        TC2_TIMED_WAIT(1, full, wph);
                        {
                            // shift16[t] = (dr * PW + ds) * pixel bytes >> 4 (host table; a 16-byte multiple in both operand modes)
                            const uint64_t boff = p.base_off_mode ? ((uint64_t)(p.ds[t] & 7) << 49) : 0ull;
                            const uint64_t pd_hi = (pd_base | boff) + (uint64_t)p.shift16[t];
                            const uint64_t pd_lo = pd_hi + (uint64_t)(p.patch_plane_bytes >> 4);
                            const uint64_t wd_lo = wd_hi + (uint64_t)(W_TILE >> 4);
                            const uint32_t first = (cb | t) ? 1u : 0u;          // 0 only for the first MMA of the tile
                            if (WIDE) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)              // A_hi x [W_hi | W_lo]: columns [0, 2 BLOCK_N)
                                    tc_mma_bf16(d_tmem, pd_hi + 2 * kk, wd_hi + 2 * kk, idesc_w, kk ? 1u : first);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(d_tmem, pd_lo + 2 * kk, wd_hi + 2 * kk, idesc, 1u);
                            } else {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    tc_mma_bf16(d_tmem, pd_hi + 2 * kk, wd_hi + 2 * kk, idesc, kk ? 1u : first);
                                if (NPASS >= 2) {
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(d_tmem, pd_lo + 2 * kk, wd_hi + 2 * kk, idesc, 1u);
                                }
                                if (NPASS >= 3) {
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(d_tmem, pd_hi + 2 * kk, wd_lo + 2 * kk, idesc, 1u);
                                }
                            }
                            tc_commit(empty);
                            if (t == p.T - 1) {
                                tc_commit(pe_bar);
                                if (cb == p.cblocks - 1) tc_commit(af(ab));
                            }
                        }
                        wd_hi += (uint64_t)(w_stage >> 4); full += 8u; empty += 8u;
                        if (++ws == p.n_ws) { ws = 0; wph ^= 1u; wd_hi = wd_first; full = wf(0); empty = we(0); }
                    }
                    if (++pb == p.n_pb) { pb = 0; pph ^= 1u; }
                }
            }
            if (dbg && leader) {
                dbg[blockIdx.x * 16 + 0] = dbg_acc[0]; dbg[blockIdx.x * 16 + 1] = dbg_acc[1];
                dbg[blockIdx.x * 16 + 2] = dbg_acc[2]; dbg[blockIdx.x * 16 + 3] = clock64() - t_start;
            }
        }
    } else {
        // ===== epilogue warps 3..6 =====
        // TMEM lane = pixel, registers = channels.  A thread's 32 channels of one pixel are staged through a per-warp 4 KB
        // shared-memory tile (16-byte granules XOR-swizzled by the row, conflict-free both ways) and leave as 512-byte
        // coalesced stores: lane l of store i writes granule (l & 7) of pixel i*4 + (l >> 3).  Bias comes from a per-warp
        // shared copy (one load per tile instead of one __ldg per element).
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int th = row >> 3, tw = row & 7;
        constexpr int CH = BLOCK_N >= 32 ? 32 : 16;
        float* s_bias = reinterpret_cast<float*>(smem_raw + (epi_base - smem_u32(smem_raw))) + q * 256;
        float4* stg = reinterpret_cast<float4*>(smem_raw + (epi_base + 4096 - smem_u32(smem_raw))) + q * 256;
        const bool vec_out = (p.Cout & 3) == 0 && CH == 32;
        const bool tanh_out = p.act == DSR_ACT_TANH;
        // Running statistics of the (sample, channel tile) this warp is in: a persistent CTA meets several tiles of one sample in
        // a row (512 tiles per 256 x 256 sample on 148 CTAs), and every warp of every tile adding to the SAME 2 * Cout fp64
        // addresses of that sample serialises in L2 - measured on the 7x7 3 -> 32 layer: 0.27 ms with the statistics against
        // 0.13 ms without, i.e. 12 samples x 2048 same-address atomics x ~6 ns.  Lane l keeps the sums of channels c0 + l of
        // both 32-column chunks in registers and adds them once per sample (RUN: BLOCK_N <= 64, at most two chunks).
        constexpr bool RUN = BLOCK_N <= 64 && CH == 32;
        double rs0 = 0.0, rq0 = 0.0, rs1 = 0.0, rq1 = 0.0;
        int run_n = -1, run_co0 = 0;
        auto run_flush = [&]() {
            if (run_n >= 0) {
                const int co_a = run_co0 + lane, co_b = run_co0 + 32 + lane;
                if (co_a < p.Cout) {
                    double* sp = stats + ((long)run_n * p.Cout + co_a) * 2;
                    atomicAdd(sp, rs0); atomicAdd(sp + 1, rq0);
                }
                if (BLOCK_N > 32 && co_b < p.Cout) {
                    double* sp = stats + ((long)run_n * p.Cout + co_b) * 2;
                    atomicAdd(sp, rs1); atomicAdd(sp + 1, rq1);
                }
            }
            rs0 = rq0 = rs1 = rq1 = 0.0;
        };
        int ti = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
            const int ab = ti & 1;
            const int phase = tile / p.tiles_per_phase;
            const int tp = tile - phase * p.tiles_per_phase;
            int r = tp / p.tiles_co;
            const int co0 = (tp - r * p.tiles_co) * BLOCK_N;
            const int tw_i = r % p.tiles_w; r /= p.tiles_w;
            const int th_i = r % p.tiles_h;
            const int n = r / p.tiles_h;
            const int h = th_i * TC2_TH + th, w = tw_i * TC2_TW + tw;
            const bool valid = (h < p.Ht) && (w < p.Wt);
            // pixel (0, 0) of the tile in the output tensor; rows of the tile are os*Wo*Cout apart, pixels os*Cout
            float* obase = out + ((((long)n * p.Ho + (long)(th_i * TC2_TH) * p.os + p.ph + (phase >> 1)) * p.Wo) +
                                  (long)(tw_i * TC2_TW) * p.os + p.pw + (phase & 1)) * p.Cout;
            const long pstride = (long)p.os * p.Cout, rstride = (long)p.os * p.Wo * p.Cout;
            float* orow = obase + (long)th * rstride + (long)tw * pstride;
#pragma unroll
            for (int k = 0; k < (BLOCK_N + 31) / 32; ++k) {
                const int co = co0 + k * 32 + lane;
                s_bias[k * 32 + lane] = (bias != nullptr && co < p.Cout) ? __ldg(bias + co) : 0.f;
            }
            __syncwarp();
            TC2_TIMED_WAIT(0, af(ab), (ti >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += CH) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * ACC + c0);
                uint32_t v2[WIDE ? 32 : 1];
                if (CH == 32) tc_ld32(taddr, v); else tc_ld16(taddr, v);
                if (WIDE) { if (CH == 32) tc_ld32(taddr + BLOCK_N, v2); else tc_ld16(taddr + BLOCK_N, v2); }
                tc_wait_ld();
                if (WIDE) {                                 // + the A_hi x W_lo half of the accumulator
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                }
                float f[32];
                const int nch = p.Cout - co0 - c0;              // channels of this chunk inside the tensor (may exceed CH)
#pragma unroll
                for (int j = 0; j < CH; ++j)
                    f[j] = (valid && j < nch) ? __uint_as_float(v[j]) * p.out_scale + s_bias[c0 + j] : 0.f;
                if (vec_out) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        if (tanh_out) { o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w); }
                        stg[lane * 8 + (j ^ (lane & 7))] = o;
                    }
                    __syncwarp();
                    const int g = lane & 7;
                    if (g * 4 < nch) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = i * 4 + (lane >> 3);            // pixel of this warp: tile row q*4 + i/2, column (i&1)*4 + lane/8
                            const int hh = q * 4 + (i >> 1), ww = (i & 1) * 4 + (lane >> 3);
                            if (th_i * TC2_TH + hh < p.Ht && tw_i * TC2_TW + ww < p.Wt) {
                                const float4 o = stg[rr * 8 + (g ^ (rr & 7))];
                                *reinterpret_cast<float4*>(obase + (long)hh * rstride + (long)ww * pstride + co0 + c0 + g * 4) = o;
                            }
                        }
                    }
                    __syncwarp();
                } else if (valid) {
#pragma unroll
                    for (int j = 0; j < CH; ++j)
                        if (j < nch) orow[co0 + c0 + j] = tanh_out ? tanhf(f[j]) : f[j];
                }
                if (stats != nullptr && CH == 32) {
                    // column sums over the warp's 32 pixels by a reduce-scatter butterfly: after the five exchange steps
                    // lane l holds the total of channel c0 + l (31 shuffles per quantity instead of 160)
                    float g2[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) g2[j] = f[j] * f[j];
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const bool up = (lane & step) != 0;
#pragma unroll
                        for (int j = 0; j < step; ++j) {
                            const float sf = up ? f[j] : f[j + step];
                            const float sg = up ? g2[j] : g2[j + step];
                            const float rf = __shfl_xor_sync(0xffffffffu, sf, step);
                            const float rg = __shfl_xor_sync(0xffffffffu, sg, step);
                            f[j] = (up ? f[j + step] : f[j]) + rf;
                            g2[j] = (up ? g2[j + step] : g2[j]) + rg;
                        }
                    }
                    if (RUN && p.run_stats) {
                        if (n != run_n || co0 != run_co0) { run_flush(); run_n = n; run_co0 = co0; }
                        if (c0 == 0) { rs0 += (double)f[0]; rq0 += (double)g2[0]; }
                        else { rs1 += (double)f[0]; rq1 += (double)g2[0]; }
                    } else {
                        const int co = co0 + c0 + lane;
                        if (co < p.Cout) {
                            double* sp = stats + ((long)n * p.Cout + co) * 2;
                            atomicAdd(sp, (double)f[0]);
                            atomicAdd(sp + 1, (double)g2[0]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ae(ab));
        }
        if (RUN && stats != nullptr) run_flush();
        if (dbg && warp == 3 && lane == 0) { dbg[blockIdx.x * 16 + 4] = dbg_acc[0]; dbg[blockIdx.x * 16 + 5] = clock64() - t_start; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#define ST(s) ((cudaStream_t)(s))
static const int TC2_SMEM_MAX = 227 * 1024;
static const int TC2_EPI_SMEM = 4096 + 16384;

template <int BLOCK_N, int NPASS>
static int launch_tc2(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
                      const Tc2Params& p, const float* bias, float* out, double* stats, int grid, int smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(conv_tc2_kernel<BLOCK_N, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_MAX) != cudaSuccess) {
            dsr_set_error("conv_tc2: cannot raise dynamic shared memory to %d", TC2_SMEM_MAX);
            return DSR_ERR_CUDA;
        }
        attr = true;
    }
    conv_tc2_kernel<BLOCK_N, NPASS><<<grid, 224, smem, st>>>(ah, al, wh, wl, p, bias, out, stats);
    return dsr_check_launch("conv_tc2");
}

template <int NPASS>
static int dispatch_tc2(int bn, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
                        const Tc2Params& p, const float* bias, float* out, double* stats, int grid, int smem, cudaStream_t st) {
    switch (bn) {
        case 16: return launch_tc2<16, NPASS>(ah, al, wh, wl, p, bias, out, stats, grid, smem, st);
        case 32: return launch_tc2<32, NPASS>(ah, al, wh, wl, p, bias, out, stats, grid, smem, st);
        case 64: return launch_tc2<64, NPASS>(ah, al, wh, wl, p, bias, out, stats, grid, smem, st);
        case 128: return launch_tc2<128, NPASS == 4 ? 3 : NPASS>(ah, al, wh, wl, p, bias, out, stats, grid, smem, st);
        case 256: return launch_tc2<256, NPASS == 4 ? 3 : NPASS>(ah, al, wh, wl, p, bias, out, stats, grid, smem, st);
    }
    dsr_set_error("conv_tc2: unsupported BLOCK_N %d", bn);
    return DSR_ERR_UNSUPPORTED;
}

extern "C" int dsr_tc2_set_debug(long long* counters) {
    if (cudaMemcpyToSymbol(g_tc2_dbg, &counters, sizeof(counters)) != cudaSuccess) {
        dsr_set_error("dsr_tc2_set_debug: cudaMemcpyToSymbol failed");
        return DSR_ERR_CUDA;
    }
    return DSR_OK;
}

static int tc2_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

extern "C" int dsr_tc_gemm2(const void* A_hi, const void* A_lo, int N, int Ha, int Wa, int Ca, const void* W_hi, const void* W_lo,
                            int Cout, int T, const int* tap_dr, const int* tap_ds, int a_off_h, int a_off_w, int Ht, int Wt,
                            const float* bias, float* out, int Ho, int Wo, int os, int ph, int pw, int nphase, int act, int npass,
                            int f16, float out_scale, double* stats, int a_mode, void* stream) {
    DSR_REQUIRE(nphase == 1 || (nphase == 4 && os == 2 && ph == 0 && pw == 0), "phases: 1, or 4 with output stride 2");
    DSR_REQUIRE(A_hi && W_hi && out && tap_dr && tap_ds, "null pointer");
    DSR_REQUIRE(a_mode == 0 || (a_mode == 1 && Ca == 8), "compact operand mode needs an 8-channel arranged tensor");
    DSR_REQUIRE(npass >= 1 && npass <= 3 && (npass < 2 || A_lo) && (npass < 3 || W_lo), "bad precision mode");
    DSR_REQUIRE(T >= 1 && T <= TC2_MAX_TAPS && ((Ca & 63) == 0 || a_mode == 1) && Cout >= 1, "bad GEMM shape");
    DSR_REQUIRE(!((uintptr_t)A_hi & 15) && !((uintptr_t)W_hi & 15) && !((uintptr_t)out & 15), "buffers must be 16-byte aligned");
    DSR_REQUIRE(!stats || act == DSR_ACT_NONE, "statistics are taken before any activation");
    Tc2Params p;
    int max_dr = 0, max_ds = 0;
    for (int t = 0; t < T; ++t) {
        if (tap_dr[t] < 0 || tap_ds[t] < 0) { dsr_set_error("conv_tc2: negative tap offset"); return DSR_ERR_UNSUPPORTED; }
        p.dr[t] = (signed char)tap_dr[t]; p.ds[t] = (signed char)tap_ds[t];
        if (tap_dr[t] > max_dr) max_dr = tap_dr[t];
        if (tap_ds[t] > max_ds) max_ds = tap_ds[t];
    }
    if (max_ds > 8 || (long)Ht * Wt < 128 || Wt < TC2_TW) {
        dsr_set_error("conv_tc2: shape not covered (tap window %d wide, output %dx%d)", max_ds + 1, Ht, Wt);
        return DSR_ERR_UNSUPPORTED;
    }
    int bn = Cout >= 256 ? 256 : (Cout > 64 ? 128 : (Cout > 32 ? 64 : (Cout > 16 ? 32 : 16)));
    if (stats && bn < 32) bn = 32;
    bn = tc2_env("DSR_TC2_BN", bn);
    const int Kt = a_mode ? 64 : Ca;          // K elements per tap in the weight matrix
    p.N = N; p.Ht = Ht; p.Wt = Wt; p.Ca = Kt; p.T = T; p.cblocks = a_mode ? 1 : Ca / 64; p.a_mode = a_mode;
    p.ah = a_off_h; p.aw = a_off_w; p.Hp = TC2_TH + max_dr;
    p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.os = os; p.ph = ph; p.pw = pw; p.act = act;
    p.f16 = f16; p.out_scale = out_scale;
    p.run_stats = tc2_env("DSR_TC2_RUNSTATS", 1);
    p.base_off_mode = tc2_env("DSR_TC2_BASEOFF", 0);   // measured on B200: the swizzle XOR uses absolute smem address bits,
                                                        // so shifted descriptor starts need NO base offset
    p.tiles_w = dsr_cdiv(Wt, TC2_TW); p.tiles_h = dsr_cdiv(Ht, TC2_TH); p.tiles_co = dsr_cdiv(Cout, bn);
    p.nphase = nphase; p.tiles_per_phase = p.tiles_w * p.tiles_h * p.tiles_co * N;
    p.total_tiles = p.tiles_per_phase * nphase;
    p.PW = tc2_env("DSR_TC2_PW", TC2_TW + max_ds + (a_mode ? 7 : 0));    // compact rows reach 7 pixels further right
    for (int t = 0; t < T; ++t) p.shift16[t] = ((unsigned)(tap_dr[t] * p.PW + tap_ds[t]) * (a_mode ? 16u : 128u)) >> 4;
    p.patch_tx_bytes = (unsigned)p.Hp * p.PW * (a_mode ? 16u : 128u);
    p.patch_plane_bytes = (p.patch_tx_bytes + 1023u) & ~1023u;
    const int na = npass >= 2 ? 2 : 1, nw = npass >= 3 ? 2 : 1;
    const long patch_set = (long)na * p.patch_plane_bytes, w_stage = (long)nw * bn * 128;
    const long budget = TC2_SMEM_MAX - 1024 - 512 - TC2_EPI_SMEM;
    // shared-memory plan: double-buffer the patch when at least 4 weight stages (>= 96 KB in flight hides the L2
    // latency at full MMA rate) still fit beside it
    const int want_ws = tc2_env("DSR_TC2_MINWS", 4);
    p.n_pb = (2 * patch_set + want_ws * w_stage <= budget && (p.cblocks > 1 || p.total_tiles > dsr_num_sms())) ? 2 : 1;
    p.n_pb = tc2_env("DSR_TC2_NPB", p.n_pb);
    long ws = (budget - p.n_pb * patch_set) / w_stage;
    if (ws > 8) ws = 8;
    if (ws < 2) { dsr_set_error("conv_tc2: tile does not fit shared memory"); return DSR_ERR_UNSUPPORTED; }
    p.n_ws = (int)ws;
    const int smem = (int)(p.n_pb * patch_set + p.n_ws * w_stage + 1024 + 512 + TC2_EPI_SMEM);

    CUtensorMap mah, mal, mwh, mwl;
    cuuint64_t adims[4] = {(cuuint64_t)Ca, (cuuint64_t)Wa, (cuuint64_t)Ha, (cuuint64_t)N};
    cuuint64_t astr[3] = {(cuuint64_t)Ca * 2, (cuuint64_t)Wa * Ca * 2, (cuuint64_t)Ha * Wa * Ca * 2};
    cuuint32_t abox[4] = {(cuuint32_t)(a_mode ? 8 : 64), (cuuint32_t)p.PW, (cuuint32_t)p.Hp, 1};
    const CUtensorMapSwizzle aswz = a_mode ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B;
    int rc = encode_map(&mah, A_hi, 4, adims, astr, abox, aswz);
    if (rc) return rc;
    mal = mah;
    if (npass >= 2 && (rc = encode_map(&mal, A_lo, 4, adims, astr, abox, aswz))) return rc;
    cuuint64_t wdims[2] = {(cuuint64_t)T * Kt, (cuuint64_t)Cout * nphase};
    cuuint64_t wstr[1] = {(cuuint64_t)T * Kt * 2};
    cuuint32_t wbox[2] = {64, (cuuint32_t)bn};
    if ((rc = encode_map(&mwh, W_hi, 2, wdims, wstr, wbox))) return rc;
    mwl = mwh;
    if (npass >= 3 && (rc = encode_map(&mwl, W_lo, 2, wdims, wstr, wbox))) return rc;
    int grid = p.total_tiles < dsr_num_sms() ? p.total_tiles : dsr_num_sms();
    if (npass == 1) return dispatch_tc2<1>(bn, mah, mal, mwh, mwl, p, bias, out, stats, grid, smem, ST(stream));
    if (npass == 2) return dispatch_tc2<2>(bn, mah, mal, mwh, mwl, p, bias, out, stats, grid, smem, ST(stream));
    if (bn <= 64 && tc2_env("DSR_TC2_WIDE", 1))      // three products in two MMAs per K step (see conv_tc2_kernel)
        return dispatch_tc2<4>(bn, mah, mal, mwh, mwl, p, bias, out, stats, grid, smem, ST(stream));
    return dispatch_tc2<3>(bn, mah, mal, mwh, mwl, p, bias, out, stats, grid, smem, ST(stream));
}
