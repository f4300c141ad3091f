// Channel-major tcgen05 implicit GEMM for layers with Cout >= 128 (third-generation convolution kernel, sm_100a).
//
// Measured on B200 (profiles/r1_*): with both operands in shared memory a 128 x 128 x 16 MMA reads
// 8 KB of smem per 64-cycle slot = the full 128 B/clk smem bandwidth, so the pixel-major kernels (conv_tc.cu,
// conv_tc2.cu) issue at ~118 cycles per MMA whenever TMA is also writing; N = 256 MMAs (12 KB per 128-cycle slot)
// run at ~95 % of the tensor floor.  This kernel therefore makes N = pixels:
//
//     D[co][pixel] += W[co][k] * A[pixel][k]          M = 128 output channels, N = 8 x TH pixels (TH <= 32 -> N <= 256)
//
//   * the accumulator is channel-major in TMEM (lane = output channel, column = pixel): the epilogue thread of a lane
//     owns ONE channel, so a warp stores 32 consecutive channels of a pixel (128-byte coalesced NHWC stores) and the
//     InstanceNorm / GroupNorm statistics (per-(n, c) sum, sum of squares) are thread-local sums - no shuffles;
//   * A is a halo-resident patch as in conv_tc2.cu, but in 32-channel K blocks (64-byte rows, SWIZZLE_64B) so that a
//     double-buffered hi+lo patch for 256 pixels and a deep weight ring fit in 227 KB together;
//   * the tile height TH is chosen per layer on the host so that the tile count fills whole waves of 148 SMs;
//   * persistent CTAs, double-buffered 256-column accumulators (all 512 TMEM columns).
// Warp roles (224 threads): 0 = patch TMA producer, 1 = weight TMA producer, 2 = TMEM allocator + MMA issuer,
// 3..6 = epilogue (TMEM lane quadrant = warp % 4).
#include "tc_common.cuh"

#define TC3_MAX_TAPS 64
#define TC3_TW 8
#define TC3_MAX_WS 12

struct Tc3Params {
    int N, Ht, Wt, TH;
    int tiles_w, tiles_h, tiles_co, total_tiles;
    int Ca, T, cblocks;       // cblocks = Ca / KB (KB = 32 or 64 channels per stage)
    int ah, aw, Hp, PW;       // patch = Hp rows x PW pixels x 64 B
    int Ho, Wo, Cout, os, ph, pw;
    int nphase, tiles_per_phase;  // 4 output phases of a stride-2 transposed conv in one launch (see conv_tc.cu)
    int act, f16;
    float out_scale;
    int n_pb, n_ws;
    unsigned patch_plane_bytes, patch_tx_bytes;
    signed char dr[TC3_MAX_TAPS], ds[TC3_MAX_TAPS];
    unsigned shift16[TC3_MAX_TAPS];   // per tap: (dr * PW + ds) * row bytes >> 4 = what the patch descriptor's address field moves by
};

// K-major SWIZZLE_64B descriptor (64-byte rows, 8-row groups): layout type 4; LBO unused (1); SBO = group stride
__device__ __forceinline__ uint64_t make_sdesc64(uint32_t saddr, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

// K-major SWIZZLE_128B descriptor (128-byte rows) with an explicit 8-row-group stride: the K = 64 form of the single-pass
// instantiation.  The swizzle XOR uses absolute shared-memory address bits (measured, conv_tc2.cu), so a patch descriptor
// may start at any pixel of a 1024-byte aligned patch without a base offset.
__device__ __forceinline__ uint64_t make_sdesc128(uint32_t saddr, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ long long* g_tc3_dbg = nullptr;
#define TC3_TIMED_WAIT(slot, bar, parity)                 \
    do {                                                  \
        if (dbg) {                                        \
            const long long t0_ = clock64();              \
            mbar_wait(bar, parity);                       \
            dbg_acc[slot] += clock64() - t0_;             \
        } else {                                          \
            mbar_wait(bar, parity);                       \
        }                                                 \
    } while (0)

// KB = K elements (channels) per pipeline stage: 32 (64-byte rows, SWIZZLE_64B: what fits beside hi + lo planes) or 64
// (128-byte rows, SWIZZLE_128B).  A TMA box is fetched row by row, ~3 cycles per row whatever its width: with ONE MMA pass
// per product a 64-byte-row stage carries only 2 MMAs (256 tensor cycles) for 128 weight rows + its share of the patch
// rows, and the single-pass kernel ran TMA-bound at 44 % of the tensor peak (ncu r2c); 128-byte rows halve the rows per
// byte (4 MMAs per stage) and the single-pass form has the shared memory for them (no lo planes).
template <int NPASS, int KB>
__global__ void __launch_bounds__(224, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
                const __grid_constant__ Tc3Params p, const float* __restrict__ bias, float* __restrict__ out,
                double* __restrict__ stats) {
    constexpr int NA = NPASS >= 2 ? 2 : 1;
    constexpr int NW = NPASS >= 3 ? 2 : 1;
    constexpr uint32_t ROWB = KB * 2;                    // bytes per operand row (one pixel / one output channel) in a stage
    constexpr uint32_t W_TILE = 128 * ROWB;              // 128 output channels x KB k x 2 B
    constexpr uint32_t ACC_COLS = 256;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t patch_set = NA * p.patch_plane_bytes;
    const uint32_t w_base = smem_base + p.n_pb * patch_set;
    const uint32_t w_stage = NW * W_TILE;
    const uint32_t bar_base = w_base + p.n_ws * w_stage;
    auto pf = [&](int i) { return bar_base + 8u * i; };
    auto pe = [&](int i) { return bar_base + 8u * (p.n_pb + i); };
    auto wf = [&](int i) { return bar_base + 8u * (2 * p.n_pb + i); };
    auto we = [&](int i) { return bar_base + 8u * (2 * p.n_pb + p.n_ws + i); };
    auto af = [&](int i) { return bar_base + 8u * (2 * p.n_pb + 2 * p.n_ws + i); };
    auto ae = [&](int i) { return bar_base + 8u * (2 * p.n_pb + 2 * p.n_ws + 2 + i); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.n_pb + 2 * p.n_ws + 4);
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* dbg = g_tc3_dbg;
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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(2u * ACC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    if (dbg && threadIdx.x == 0) dbg[blockIdx.x * 16 + 9] = clock64() - t_start;     // set-up time

    if (warp == 0) {
        // ===== patch producer: one box per (tile, 32-channel block) =====
        // (ring positions and barrier parities advance incrementally in all three roles: a runtime `i % n` / `i / n` pair is a
        // ~30-instruction dependent chain through MUFU.RCP, and the MMA issuer's loop was LATENCY-bound at ~860 cycles per
        // stage - ncu source view r2z - against 512 cycles of tensor work)
        {
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
                const int hc = th_i * p.TH + p.ah + (phase >> 1), wc = tw_i * TC3_TW + p.aw + (phase & 1);
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    TC3_TIMED_WAIT(0, pe(pb), pph);
                    {
                        const uint32_t dst = smem_base + pb * patch_set;
                        mbar_expect_tx(pf(pb), NA * p.patch_tx_bytes);
                        tma_load_4d(dst, &mapA_hi, pf(pb), cb * KB, wc, hc, n);
                        if (NPASS >= 2) tma_load_4d(dst + p.patch_plane_bytes, &mapA_lo, pf(pb), cb * KB, wc, hc, n);
                    }
                    if (++pb == p.n_pb) { pb = 0; pph ^= 1u; }
                }
            }
            if (dbg && leader) dbg[blockIdx.x * 16 + 6] = dbg_acc[0];
        }
    } else if (warp == 1) {
        // ===== weight producer: one [128 co x 32 k] tile (hi, lo) per (tile, block, tap) =====
        {
            int ws = 0;
            uint32_t wph = 1u, dst = w_base, full = wf(0), empty = we(0);
            const bool leader = elect_one();
            if (leader)
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int phase = tile / p.tiles_per_phase;
                const int co0 = ((tile - phase * p.tiles_per_phase) % p.tiles_co) * 128 + phase * p.Cout;
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    int kw = cb * KB;
                    for (int t = 0; t < p.T; ++t, kw += p.Ca) {
                        TC3_TIMED_WAIT(0, empty, wph);
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
        // ===== MMA issuer: A operand = weights (M = 128 channels), B operand = shifted patch (N = 8*TH pixels).
        // The whole warp runs the (warp-uniform) loop and the waits; one elected lane issues. =====
        {
            const uint32_t idesc = make_idesc(128, 8 * p.TH, p.f16 ? 0u : 1u);
            const uint64_t wdesc0 = KB == 32 ? make_sdesc64(0, 512) : make_sdesc128(0, 1024);   // address field added per MMA
            const uint64_t pdesc0 = KB == 32 ? make_sdesc64(0, (uint32_t)p.PW * 64u) : make_sdesc128(0, (uint32_t)p.PW * 128u);
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
                TC3_TIMED_WAIT(2, ae(ab), ((ti >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)ab * ACC_COLS;
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    TC3_TIMED_WAIT(0, pf(pb), pph);
                    const uint32_t patch_hi = smem_base + pb * patch_set;
                    const uint64_t pd_base = pdesc0 + (uint64_t)((patch_hi & 0x3FFFF) >> 4);
                    const uint32_t pe_bar = pe(pb);
                    for (int t = 0; t < p.T; ++t) {
                        TC3_TIMED_WAIT(1, full, wph);
                        {
                            const uint64_t wd_lo = wd_hi + (uint64_t)(W_TILE >> 4);
                            const uint64_t pd_hi = pd_base + (uint64_t)p.shift16[t];
                            const uint64_t pd_lo = pd_hi + (uint64_t)(p.patch_plane_bytes >> 4);
                            const uint32_t first = (cb | t) ? 1u : 0u;          // 0 only for the first MMA of the tile
#pragma unroll
                            for (int kk = 0; kk < KB / 16; ++kk)                // one K = 16 MMA per 32 bytes of the rows
                                tc_mma_bf16(d_tmem, wd_hi + 2 * kk, pd_hi + 2 * kk, idesc, kk ? 1u : first);
                            if (NPASS >= 2) {
#pragma unroll
                                for (int kk = 0; kk < KB / 16; ++kk) tc_mma_bf16(d_tmem, wd_hi + 2 * kk, pd_lo + 2 * kk, idesc, 1u);
                            }
                            if (NPASS >= 3) {
#pragma unroll
                                for (int kk = 0; kk < KB / 16; ++kk) tc_mma_bf16(d_tmem, wd_lo + 2 * kk, pd_hi + 2 * kk, idesc, 1u);
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
        // ===== epilogue warps 3..6: lane = output channel, columns = pixels =====
        const int q = warp & 3;
        const int ncols = 8 * p.TH;
        const long pstride = (long)p.os * p.Cout;                 // between horizontally adjacent output pixels
        const long rstride = (long)p.os * p.Wo * p.Cout;          // between tile rows
        const bool tanh_out = p.act == DSR_ACT_TANH;
        int ti = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
            const int ab = ti & 1;
            const int phase = tile / p.tiles_per_phase;
            const int tp = tile - phase * p.tiles_per_phase;
            int r = tp / p.tiles_co;
            const int co = (tp - r * p.tiles_co) * 128 + q * 32 + lane;
            const int tw_i = r % p.tiles_w; r /= p.tiles_w;
            const int th_i = r % p.tiles_h;
            const int n = r / p.tiles_h;
            const int h0 = th_i * p.TH, w0 = tw_i * TC3_TW;
            const bool cvalid = co < p.Cout;
            const float bv = (bias != nullptr && cvalid) ? __ldg(bias + co) : 0.f;
            const int wvalid = p.Wt - w0 < 8 ? p.Wt - w0 : 8;     // pixels of a tile row inside the image
            int hvalid = p.Ht - h0;                                // rows of the tile inside the image
            if (hvalid > p.TH) hvalid = p.TH;
            if (!cvalid) hvalid = 0;
            float* obase = out + (((long)n * p.Ho + (long)h0 * p.os + p.ph + (phase >> 1)) * p.Wo + (long)w0 * p.os + p.pw + (phase & 1)) * p.Cout + co;
            float ssum = 0.f, ssq = 0.f;
            TC3_TIMED_WAIT(0, af(ab), (ti >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < ncols; c0 += 16) {
                uint32_t v[16];
                const long long tl0 = dbg ? clock64() : 0;
                tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * ACC_COLS + c0), v);
                tc_wait_ld();
                if (dbg) dbg_acc[1] += clock64() - tl0;
                // 16 columns = 2 tile rows of 8 pixels
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int hr = (c0 >> 3) + hh;
                    if (hr < hvalid) {
                        float* orow = obase + (long)hr * rstride;
                        float x[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[hh * 8 + j]) * p.out_scale + bv;
                        if (wvalid == 8) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) { ssum += x[j]; ssq += x[j] * x[j]; }
                            if (tanh_out) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) x[j] = tanhf(x[j]);
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) orow[j * pstride] = x[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (j < wvalid) {
                                    ssum += x[j]; ssq += x[j] * x[j];
                                    orow[j * pstride] = tanh_out ? tanhf(x[j]) : x[j];
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ae(ab));
            if (stats != nullptr && cvalid) {
                double* sp = stats + ((long)n * p.Cout + co) * 2;
                atomicAdd(sp, (double)ssum);
                atomicAdd(sp + 1, (double)ssq);
            }
        }
        if (dbg && warp == 3 && lane == 0) {
            dbg[blockIdx.x * 16 + 4] = dbg_acc[0]; dbg[blockIdx.x * 16 + 5] = clock64() - t_start;
            dbg[blockIdx.x * 16 + 8] = dbg_acc[1];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * ACC_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#define ST(s) ((cudaStream_t)(s))
static const int TC3_SMEM_MAX = 227 * 1024;

extern "C" int dsr_tc3_set_debug(long long* counters) {
    if (cudaMemcpyToSymbol(g_tc3_dbg, &counters, sizeof(counters)) != cudaSuccess) {
        dsr_set_error("dsr_tc3_set_debug: cudaMemcpyToSymbol failed");
        return DSR_ERR_CUDA;
    }
    return DSR_OK;
}

static int tc3_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

static int encode_map64(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                        const cuuint32_t* box, bool sw128 = false) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { dsr_set_error("conv_tc3: cuTensorMapEncodeTiled entry point unavailable"); return DSR_ERR_CUDA; }
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dsr_set_error("conv_tc3: cuTensorMapEncodeTiled failed (%d)", (int)r); return DSR_ERR_CUDA; }
    return DSR_OK;
}

template <int NPASS, int KB>
static int launch_tc3(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
                      const Tc3Params& p, const float* bias, float* out, double* stats, int grid, int smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(conv_tc3_kernel<NPASS, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM_MAX) != cudaSuccess) {
            dsr_set_error("conv_tc3: cannot raise dynamic shared memory to %d", TC3_SMEM_MAX);
            return DSR_ERR_CUDA;
        }
        attr = true;
    }
    conv_tc3_kernel<NPASS, KB><<<grid, 224, smem, st>>>(ah, al, wh, wl, p, bias, out, stats);
    return dsr_check_launch("conv_tc3");
}

// tile height that fills whole waves of SMs: minimise rounds x tile cost (MMA time ~ TH, plus a fixed per-tile cost)
static int tc3_pick_th(int N, int Ht, int Wt, int tiles_co, int sms) {
    int best = 16;
    double best_cost = 1e30;
    for (int th = 16; th <= 32; th += 2) {
        const long tiles = (long)N * dsr_cdiv(Wt, TC3_TW) * dsr_cdiv(Ht, th) * tiles_co;
        const long rounds = (tiles + sms - 1) / sms;
        // N = 128 MMAs run at ~55 % of the tensor floor (smem-read bound), N = 256 at ~95 %: model that as a slope
        const double eff = 0.55 + 0.40 * (th - 16) / 16.0;
        const double cost = rounds * (th / eff + 1.5);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = th; }
    }
    return best;
}

extern "C" int dsr_tc_gemm3(const void* A_hi, const void* A_lo, int N, int Ha, int Wa, int Ca, const void* W_hi, const void* W_lo,
                            int Cout, int T, const int* tap_dr, const int* tap_ds, int a_off_h, int a_off_w, int Ht, int Wt,
                            const float* bias, float* out, int Ho, int Wo, int os, int ph, int pw, int nphase, int act, int npass,
                            int f16, float out_scale, double* stats, void* stream) {
    DSR_REQUIRE(nphase == 1 || (nphase == 4 && os == 2 && ph == 0 && pw == 0), "phases: 1, or 4 with output stride 2");
    DSR_REQUIRE(A_hi && W_hi && out && tap_dr && tap_ds, "null pointer");
    DSR_REQUIRE(npass >= 1 && npass <= 3 && (npass < 2 || A_lo) && (npass < 3 || W_lo), "bad precision mode");
    DSR_REQUIRE(T >= 1 && T <= TC3_MAX_TAPS && (Ca & 63) == 0 && Cout >= 1, "bad GEMM shape");
    DSR_REQUIRE(!((uintptr_t)A_hi & 15) && !((uintptr_t)W_hi & 15) && !((uintptr_t)out & 15), "buffers must be 16-byte aligned");
    DSR_REQUIRE(!stats || act == DSR_ACT_NONE, "statistics are taken before any activation");
    Tc3Params p;
    int max_dr = 0, max_ds = 0;
    for (int t = 0; t < T; ++t) {
        if (tap_dr[t] < 0 || tap_ds[t] < 0) { dsr_set_error("conv_tc3: negative tap offset"); return DSR_ERR_UNSUPPORTED; }
        p.dr[t] = (signed char)tap_dr[t]; p.ds[t] = (signed char)tap_ds[t];
        if (tap_dr[t] > max_dr) max_dr = tap_dr[t];
        if (tap_ds[t] > max_ds) max_ds = tap_ds[t];
    }
    if (max_ds > 8 || max_dr > 8 || Ht < 16 || Wt < TC3_TW) {
        dsr_set_error("conv_tc3: shape not covered (tap window %dx%d, output %dx%d)", max_dr + 1, max_ds + 1, Ht, Wt);
        return DSR_ERR_UNSUPPORTED;
    }
    // single pass: 64-channel stages (128-byte rows) - see the kernel's header; DSR_TC3_KB=32 keeps the 64-byte rows (A/B)
    const int KB = (npass == 1 && tc3_env("DSR_TC3_KB", 64) == 64) ? 64 : 32;
    const unsigned rowb = (unsigned)KB * 2u;
    p.N = N; p.Ht = Ht; p.Wt = Wt; p.Ca = Ca; p.T = T; p.cblocks = Ca / KB;
    p.tiles_co = dsr_cdiv(Cout, 128);
    p.TH = tc3_env("DSR_TC3_TH", tc3_pick_th(N, Ht, Wt, p.tiles_co * nphase, dsr_num_sms()));
    if (p.TH < 2 || p.TH > 32 || (p.TH & 1)) { dsr_set_error("conv_tc3: bad tile height %d", p.TH); return DSR_ERR_ARG; }
    p.ah = a_off_h; p.aw = a_off_w; p.Hp = p.TH + max_dr; p.PW = TC3_TW + max_ds;
    p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.os = os; p.ph = ph; p.pw = pw; p.act = act;
    p.f16 = f16; p.out_scale = out_scale;
    p.tiles_w = dsr_cdiv(Wt, TC3_TW); p.tiles_h = dsr_cdiv(Ht, p.TH);
    p.nphase = nphase; p.tiles_per_phase = p.tiles_w * p.tiles_h * p.tiles_co * N;
    p.total_tiles = p.tiles_per_phase * nphase;
    for (int t = 0; t < T; ++t) p.shift16[t] = ((unsigned)(tap_dr[t] * p.PW + tap_ds[t]) * rowb) >> 4;
    p.patch_tx_bytes = (unsigned)p.Hp * p.PW * rowb;
    p.patch_plane_bytes = (p.patch_tx_bytes + 1023u) & ~1023u;
    const int na = npass >= 2 ? 2 : 1, nw = npass >= 3 ? 2 : 1;
    const long patch_set = (long)na * p.patch_plane_bytes, w_stage = (long)nw * 128 * rowb;
    const long budget = TC3_SMEM_MAX - 1024 - 512;
    p.n_pb = tc3_env("DSR_TC3_NPB", 2);
    long ws = (budget - p.n_pb * patch_set) / w_stage;
    if (ws > TC3_MAX_WS) ws = TC3_MAX_WS;
    ws = tc3_env("DSR_TC3_NWS", (int)ws);
    if (ws < 2) { dsr_set_error("conv_tc3: tile does not fit shared memory"); return DSR_ERR_UNSUPPORTED; }
    p.n_ws = (int)ws;
    const int smem = (int)(p.n_pb * patch_set + p.n_ws * w_stage + 1024 + 512);

    CUtensorMap mah, mal, mwh, mwl;
    cuuint64_t adims[4] = {(cuuint64_t)Ca, (cuuint64_t)Wa, (cuuint64_t)Ha, (cuuint64_t)N};
    cuuint64_t astr[3] = {(cuuint64_t)Ca * 2, (cuuint64_t)Wa * Ca * 2, (cuuint64_t)Ha * Wa * Ca * 2};
    cuuint32_t abox[4] = {(cuuint32_t)KB, (cuuint32_t)p.PW, (cuuint32_t)p.Hp, 1};
    int rc = encode_map64(&mah, A_hi, 4, adims, astr, abox, KB == 64);
    if (rc) return rc;
    mal = mah;
    if (npass >= 2 && (rc = encode_map64(&mal, A_lo, 4, adims, astr, abox))) return rc;
    cuuint64_t wdims[2] = {(cuuint64_t)T * Ca, (cuuint64_t)Cout * nphase};
    cuuint64_t wstr[1] = {(cuuint64_t)T * Ca * 2};
    cuuint32_t wbox[2] = {(cuuint32_t)KB, 128};
    if ((rc = encode_map64(&mwh, W_hi, 2, wdims, wstr, wbox, KB == 64))) return rc;
    mwl = mwh;
    if (npass >= 3 && (rc = encode_map64(&mwl, W_lo, 2, wdims, wstr, wbox))) return rc;
    int grid = p.total_tiles < dsr_num_sms() ? p.total_tiles : dsr_num_sms();
    if (npass == 1 && KB == 64) return launch_tc3<1, 64>(mah, mal, mwh, mwl, p, bias, out, stats, grid, smem, ST(stream));
    if (npass == 1) return launch_tc3<1, 32>(mah, mal, mwh, mwl, p, bias, out, stats, grid, smem, ST(stream));
    if (npass == 2) return launch_tc3<2, 32>(mah, mal, mwh, mwl, p, bias, out, stats, grid, smem, ST(stream));
    return launch_tc3<3, 32>(mah, mal, mwh, mwl, p, bias, out, stats, grid, smem, ST(stream));
}
