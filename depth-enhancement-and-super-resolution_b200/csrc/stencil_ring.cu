// Row-ring forms of the edge-aware smoothness level (main_model.py:22-73) for large plane sets (full-size frames, big batches).
//
// (Measured, r82-r85: the same ring for the total-variation sum - one plane per sample, ~12 instructions of work per pixel -
// reaches 60 % of the HBM peak against 66 % for its register kernel, so TV stays there; with 640 column quads per group the
// consumers are bound by their own latency chain: 20 warps 68 %, 10 warps 64 %, 5 warps 49 %.)
//
// The register kernels in stencil_tiled.cu are bound by load latency (ncu: long-scoreboard stalls, ~50 % of the HBM peak):
// a thread can only keep the loads of its own 4 x 4 pixels in flight.  Here the bytes in flight are decoupled from the
// threads: one producer warp streams whole image rows (all planes of a sample: depth + C image channels) into a ring of
// shared-memory row slots with 1-D bulk copies (cp.async.bulk, one copy per plane and "group" of 1..16 consecutive rows -
// contiguous in every plane - completion counted on the group's mbarrier), up to ~160 KB in flight per SM; the consumer warps compute from shared memory (16-byte reads) and hand
// the slots back through `empty` mbarriers.  The image rows are split evenly over one CTA per SM; a CTA reads every row of its
// share once (+ one or two halo rows per sample it touches) and keeps the pipeline full across samples.
#include "tc_common.cuh"
#include <stdlib.h>

#define RING_MAX_WARPS 20

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float sgnw(float w, float x) { return x != 0.f ? copysignf(w, x) : 0.f; }        // w * sign(x), w > 0

struct RingGeom {
    int B, h, w, G, NGS;
};
// Work split: the B * h image rows are divided evenly over the CTAs (a CTA's share may span several samples); a "segment" is
// the part of the share inside one sample: computed rows [r0, r1), loaded rows [lo, hi] (one halo row below; above too for
// the backward pass), in `ng` groups of G rows.
struct RingSeg {
    long gr, g1;
    __device__ __forceinline__ RingSeg(const RingGeom& g) {
        const long tot = (long)g.B * g.h;
        gr = tot * blockIdx.x / gridDim.x;
        g1 = tot * (blockIdx.x + 1) / gridDim.x;
    }
    __device__ __forceinline__ bool more() const { return gr < g1; }
    template <bool BWD>
    __device__ __forceinline__ void get(const RingGeom& g, int& b, int& r0, int& r1, int& lo, int& hi, int& ng) {
        b = (int)(gr / g.h);
        r0 = (int)(gr - (long)b * g.h);
        r1 = (int)min((long)g.h, r0 + (g1 - gr));
        lo = BWD ? max(r0 - 1, 0) : r0;
        hi = min(r1, g.h - 1);
        ng = (hi - lo + g.G) / g.G;
        gr += r1 - r0;
    }
};

template <int C, bool BWD>
__global__ void __launch_bounds__(32 * (RING_MAX_WARPS + 1))
smooth_ring_kernel(const float* __restrict__ d, const float* __restrict__ img, RingGeom g, const float* __restrict__ gscale,
                   float cx, float cy, float* __restrict__ gd, int accumulate, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char ring_raw[];
    __shared__ double red[32];
    float* ring = reinterpret_cast<float*>(ring_raw);
    const int w = g.w, h = g.h, G = g.G, NGS = g.NGS;
    const int rowf = (C + 1) * w, pstr = G * w;                    // group slot s, plane p, row rr at ring + ((s * (C+1) + p) * G + rr) * w
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)G * NGS * rowf);       // full[NGS], empty[NGS]
    const int ncw = (blockDim.x >> 5) - 1;                         // consumer warps; the last warp is the producer
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long plane = (long)h * w;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NGS; ++s) { mbar_init(smem_u32(bars + s), 1); mbar_init(smem_u32(bars + NGS + s), ncw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float fx = 0.f, fy = 0.f;
    if (warp == ncw) {
        // ---------------- producer ----------------
        // the rows of one group are contiguous in every plane: ONE bulk copy per plane and group (lanes 0..C issue them)
        int q = 0;
        for (RingSeg seg(g); seg.more();) {
            int b, r0, r1, lo, hi, ng;
            seg.template get<BWD>(g, b, r0, r1, lo, hi, ng);
            for (int k = 0; k < ng; ++k, ++q) {
                const int s = q % NGS, use = q / NGS;
                const int ra = lo + k * G, nrows = min(G, hi - ra + 1);
                const uint32_t fb = smem_u32(bars + s);
                if (lane == 0) {
                    if (use > 0) mbar_wait(smem_u32(bars + NGS + s), (use - 1) & 1);
                    mbar_expect_tx(fb, (uint32_t)nrows * rowf * 4u);
                }
                __syncwarp();
                if (lane <= C) {
                    const float* src = (lane == 0 ? d + b * plane : img + ((long)b * C + lane - 1) * plane) + (long)ra * w;
                    bulk_g2s(smem_u32(ring + ((size_t)s * (C + 1) + lane) * G * w), src, (uint32_t)nrows * w * 4u, fb);
                }
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int nct = ncw * 32, ctid = threadIdx.x;
        const int wq = w >> 2;
        const int rr0 = ctid / wq, jq0 = ctid - rr0 * wq, rr_step = nct / wq, jq_step = nct - rr_step * wq;
        const float invC = 1.f / (float)C;
        const float gs = BWD ? (gscale ? *gscale : 1.f) : 0.f;
        int qbase = 0;
        for (RingSeg seg(g); seg.more();) {
            int b, r0, r1, lo, hi, ng;
            seg.template get<BWD>(g, b, r0, r1, lo, hi, ng);
            for (int k = 0; k < ng; ++k) {
                const int q = qbase + k, s = q % NGS;
                // one lane per warp polls (640 threads polling one mbarrier word serialise in the shared-memory pipe: measured
                // ~1.3 us per group whatever its size); __syncwarp hands the observation to the other lanes
                if (lane == 0) {
                    if (k == 0) mbar_wait(smem_u32(bars + s), (q / NGS) & 1);        // later groups were waited for as "next"
                    if (k + 1 < ng) mbar_wait(smem_u32(bars + (q + 1) % NGS), ((q + 1) / NGS) & 1);
                }
                __syncwarp();
                const int ra = lo + k * G;
                const int ia = max(ra, r0), ib = min(min(ra + G, hi + 1), r1);          // rows of this group that are computed
                const int nrow_c = ib - ia;
                // slot of absolute row r (in group k - 1, k or k + 1 of this chunk): compares instead of divisions
                const int s_prev = (q + NGS - 1) % NGS, s_next = (q + 1) % NGS;
                auto slot = [&](int r) {                  // float offset of the depth plane's row r; image plane ch at + ch * pstr
                    const int o = r - ra;
                    return o < 0 ? (s_prev * (C + 1) * G + o + G) * w : (o < G ? (s * (C + 1) * G + o) * w : (s_next * (C + 1) * G + o - G) * w);
                };
                // items (row rr, column quad jq) walked with the thread's fixed stride: no per-item division
                for (int rr = rr0, jq = jq0; rr < nrow_c; ) {
                    const int j = jq << 2;
                    const int i = ia + rr;
                    jq += jq_step; rr += rr_step;
                    if (jq >= wq) { jq -= wq; ++rr; }
                    const float* pc = ring + slot(i) + j;
                    const float* pl = ring + slot(min(i + 1, h - 1)) + j;
                    const bool has_r = j + 4 < w;
                    if (!BWD) {
                        float sx[4] = {0.f, 0.f, 0.f, 0.f}, sy[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int ch = 1; ch <= C; ++ch) {
                            const float4 a = ld4(pc + ch * pstr), l = ld4(pl + ch * pstr);
                            const float nx = has_r ? pc[ch * pstr + 4] : a.w;
                            sx[0] += fabsf(a.x - l.x); sx[1] += fabsf(a.y - l.y); sx[2] += fabsf(a.z - l.z); sx[3] += fabsf(a.w - l.w);
                            sy[0] += fabsf(a.x - a.y); sy[1] += fabsf(a.y - a.z); sy[2] += fabsf(a.z - a.w); sy[3] += fabsf(a.w - nx);
                        }
                        const float4 a = ld4(pc), l = ld4(pl);
                        const float nx = has_r ? pc[4] : a.w;
                        const float cc[5] = {a.x, a.y, a.z, a.w, nx}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            fx += fabsf((cc[e] - ll[e]) * __expf(-sx[e] * invC));
                            fy += fabsf((cc[e] - cc[e + 1]) * __expf(-sy[e] * invC));
                        }
                    } else {
                        const float* pu = ring + slot(max(i - 1, 0)) + j;
                        const bool has_l = j > 0;
                        float su[4] = {0.f, 0.f, 0.f, 0.f}, sl[4] = {0.f, 0.f, 0.f, 0.f}, sh[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int ch = 1; ch <= C; ++ch) {
                            const float4 u = ld4(pu + ch * pstr), a = ld4(pc + ch * pstr), l = ld4(pl + ch * pstr);
                            const float px = has_l ? pc[ch * pstr - 1] : a.x, nx = has_r ? pc[ch * pstr + 4] : a.w;
                            su[0] += fabsf(u.x - a.x); su[1] += fabsf(u.y - a.y); su[2] += fabsf(u.z - a.z); su[3] += fabsf(u.w - a.w);
                            sl[0] += fabsf(a.x - l.x); sl[1] += fabsf(a.y - l.y); sl[2] += fabsf(a.z - l.z); sl[3] += fabsf(a.w - l.w);
                            sh[0] += fabsf(px - a.x); sh[1] += fabsf(a.x - a.y); sh[2] += fabsf(a.y - a.z);
                            sh[3] += fabsf(a.z - a.w); sh[4] += fabsf(a.w - nx);
                        }
                        const float4 u4 = ld4(pu), a = ld4(pc), l4 = ld4(pl);
                        const float px = has_l ? pc[-1] : a.x, nx = has_r ? pc[4] : a.w;
                        const float cc[6] = {px, a.x, a.y, a.z, a.w, nx}, uu[4] = {u4.x, u4.y, u4.z, u4.w}, ll[4] = {l4.x, l4.y, l4.z, l4.w};
                        float wh[5], o[4];
#pragma unroll
                        for (int e = 0; e < 5; ++e) wh[e] = __expf(-sh[e] * invC);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float v = cc[e + 1];
                            const float wl = __expf(-sl[e] * invC), wu = __expf(-su[e] * invC);
                            // d|x w|/dx = w sign(x w) = w sign(x) (w = exp(.) > 0): the weight with the sign of the difference
                            const float tv = sgnw(wl, v - ll[e]) - sgnw(wu, uu[e] - v);
                            const float th = sgnw(wh[e + 1], v - cc[e + 2]) - sgnw(wh[e], cc[e] - v);
                            o[e] = gs * (cx * tv + cy * th);
                        }
                        float* op = gd + b * plane + (long)i * w + j;
                        if (accumulate) {
                            const float4 old = ld4(op);
                            st4(op, make_float4(old.x + o[0], old.y + o[1], old.z + o[2], old.w + o[3]));
                        } else {
                            st4(op, make_float4(o[0], o[1], o[2], o[3]));
                        }
                    }
                }
                // hand slots back: forward frees this group; backward keeps it as the next group's upper halo
                __syncwarp();
                if (lane == 0) {
                    if (!BWD) mbar_arrive(smem_u32(bars + NGS + s));
                    else {
                        if (k > 0) mbar_arrive(smem_u32(bars + NGS + (q - 1) % NGS));
                        if (k == ng - 1) mbar_arrive(smem_u32(bars + NGS + s));
                    }
                }
            }
            qbase += ng;
        }
    }
    if (!BWD) {
        const double ax = block_sum<double>((double)fx, red), ay = block_sum<double>((double)fy, red);
        if (threadIdx.x == 0) { atomicAdd(out, ax); atomicAdd(out + 1, ay); }
    }
}

// Picks the ring geometry for (B, C, h, w); returns 0 when the shape does not suit the ring (rows that do not fit, tiny inputs)
static int ring_plan(int B, int C, int h, int w, RingGeom* g, int* nthreads, size_t* smem, int* grid) {
    if ((w & 3) || w < 64) return 0;
    const size_t rowbytes = (size_t)(C + 1) * w * 4;
    int G = (int)(40960 / rowbytes);
    if (G < 1) return 0;
    if (G > 16) G = 16;                     // narrow / single-plane rows: more rows per barrier, so the per-group bookkeeping amortises
    if (G > h) G = h;
    const size_t budget = 200 * 1024;
    int NGS = (int)(budget / (G * rowbytes));
    if (NGS > 8) NGS = 8;
    if (NGS < 4) return 0;
    // consumer warps: the count in 8..20 that wastes the fewest lanes on a full group
    const int items = G * (w / 4);
    int best = 8; double beff = 0.0;
    for (int cw = 8; cw <= RING_MAX_WARPS; ++cw) {
        const int T = cw * 32, rounds = (items + T - 1) / T;
        const double eff = (double)items / ((double)rounds * T);
        if (eff > beff + 1e-9 || (eff > beff - 1e-9 && cw > best)) { beff = eff; best = cw; }
    }
    if (const char* e = getenv("DSR_RING_WARPS")) {           // tuning knob (scripts/bench_stencils.py sweeps it)
        const int v = atoi(e);
        if (v >= 1 && v <= RING_MAX_WARPS) best = v;
    }
    const int sms = dsr_num_sms();
    const long share = ((long)B * h + 4L * G - 1) / (4L * G);            // at least ~4 groups of rows per CTA
    g->B = B; g->h = h; g->w = w; g->G = G; g->NGS = NGS;
    *nthreads = 32 * (best + 1);
    *smem = (size_t)G * NGS * rowbytes + 2 * (size_t)NGS * sizeof(uint64_t);
    const long ctas = (long)sms * (*smem <= 100 * 1024 ? 2 : 1);         // small rings: two CTAs per SM
    *grid = (int)(share < ctas ? share : ctas);
    return 1;
}

template <int C, bool BWD>
static int ring_launch(const float* d, const float* img, const RingGeom& g, int nthreads, size_t smem, int grid, const float* gscale,
                       float cx, float cy, float* gd, int accumulate, double* out, cudaStream_t st) {
    static size_t raised = 0;
    if (smem > raised) {
        if (cudaFuncSetAttribute(smooth_ring_kernel<C, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            dsr_set_error("smooth ring: cannot raise dynamic shared memory to %d", (int)smem);
            return DSR_ERR_CUDA;
        }
        raised = smem;
    }
    smooth_ring_kernel<C, BWD><<<grid, nthreads, smem, st>>>(d, img, g, gscale, cx, cy, gd, accumulate, out);
    return dsr_check_launch(BWD ? "smooth_level_bwd (ring)" : "smooth_level_fwd (ring)");
}

extern "C" int dsr_smooth_level_fwd_ring(const float* d, const float* img, int B, int C, int h, int w, double* out2, void* stream) {
    DSR_REQUIRE(d && img && out2 && B > 0 && C >= 1 && C <= 4 && h > 0 && w > 0, "bad arguments (C <= 4)");
    DSR_REQUIRE(!((uintptr_t)d & 15) && !((uintptr_t)img & 15), "16-byte aligned planes");
    RingGeom g; int nt, grid; size_t smem;
    DSR_REQUIRE(ring_plan(B, C, h, w, &g, &nt, &smem, &grid), "shape does not suit the row ring (w % 4 == 0, w >= 64, rows must fit shared memory)");
    switch (C) {
        case 1: return ring_launch<1, false>(d, img, g, nt, smem, grid, nullptr, 0.f, 0.f, nullptr, 0, out2, ST(stream));
        case 2: return ring_launch<2, false>(d, img, g, nt, smem, grid, nullptr, 0.f, 0.f, nullptr, 0, out2, ST(stream));
        case 3: return ring_launch<3, false>(d, img, g, nt, smem, grid, nullptr, 0.f, 0.f, nullptr, 0, out2, ST(stream));
        default: return ring_launch<4, false>(d, img, g, nt, smem, grid, nullptr, 0.f, 0.f, nullptr, 0, out2, ST(stream));
    }
}
extern "C" int dsr_smooth_level_bwd_ring(const float* d, const float* img, int B, int C, int h, int w, const float* gscale,
                                         float cx, float cy, float* gd, int accumulate, void* stream) {
    DSR_REQUIRE(d && img && gd && B > 0 && C >= 1 && C <= 4 && h > 0 && w > 0, "bad arguments (C <= 4)");
    DSR_REQUIRE(!((uintptr_t)d & 15) && !((uintptr_t)img & 15) && !((uintptr_t)gd & 15), "16-byte aligned planes");
    RingGeom g; int nt, grid; size_t smem;
    DSR_REQUIRE(ring_plan(B, C, h, w, &g, &nt, &smem, &grid), "shape does not suit the row ring (w % 4 == 0, w >= 64, rows must fit shared memory)");
    switch (C) {
        case 1: return ring_launch<1, true>(d, img, g, nt, smem, grid, gscale, cx, cy, gd, accumulate, nullptr, ST(stream));
        case 2: return ring_launch<2, true>(d, img, g, nt, smem, grid, gscale, cx, cy, gd, accumulate, nullptr, ST(stream));
        case 3: return ring_launch<3, true>(d, img, g, nt, smem, grid, gscale, cx, cy, gd, accumulate, nullptr, ST(stream));
        default: return ring_launch<4, true>(d, img, g, nt, smem, grid, gscale, cx, cy, gd, accumulate, nullptr, ST(stream));
    }
}
// used by dsr_smooth_level_fwd / _bwd: 1 when the plane set is large enough for the ring to pay (>= 2 M pixels) and fits
extern "C" int dsr_smooth_ring_suits(int B, int C, int h, int w) {
    RingGeom g; int nt, grid; size_t smem;
    return (long)B * h * w >= (2L << 20) && ring_plan(B, C, h, w, &g, &nt, &smem, &grid);
}

