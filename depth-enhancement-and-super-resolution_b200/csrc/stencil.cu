// Remaining untiled kernels of the loss stack (monitoring sums, bilinear pyramid resize, SSIM); the depth-derived
// stencils proper live in stencil_tiled.cu.  (HBM-bound; NCHW fp32 planes exactly as the reference holds them.)  Reference call sites: models/main_model.py:208-230 (masks),
// :257-298 (rectangle holes), :340-417 (loss stack), models/norms.py (normals),
// models/pytorch_ssim/__init__.py (SSIM).  C-ABI entry points at the bottom (include/dsr_b200.h).
#include <stdlib.h>
#include "common.cuh"
#include "stencil_math.cuh"
#include "../../include/dsr_b200.h"

#define TPB 256

// sums for the monitoring scalars (main_model.py:308-318): out = [sum d*m, sum p*m, sum |d*m - p*m|]
__global__ void masked_sums_kernel(const float* __restrict__ d, const float* __restrict__ p,
                                   const float* __restrict__ m, long total, double* __restrict__ out) {
    __shared__ double red[32];
    double s0 = 0, s1 = 0, s2 = 0;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        float mm = m[idx], x = d[idx] * mm, y = p[idx] * mm;
        s0 += (double)x; s1 += (double)y; s2 += (double)fabsf(x - y);
    }
    s0 = block_sum<double>(s0, red); s1 = block_sum<double>(s1, red); s2 = block_sum<double>(s2, red);
    if (threadIdx.x == 0) { atomicAdd(out, s0); atomicAdd(out + 1, s1); atomicAdd(out + 2, s2); }
}

// ------------------------------------------------------------------------------------------
// smoothness (main_model.py:22-73)
// ------------------------------------------------------------------------------------------
// bilinear resize, align_corners=True, of (BC, H, W) planes to (BC, nh, nw)
__global__ void bilinear_ac_fwd_kernel(const float* __restrict__ x, long planes, int H, int W, int nh, int nw,
                                       float* __restrict__ out) {
    long total = planes * nh * nw;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int ox = (int)(idx % nw);
        int oy = (int)((idx / nw) % nh);
        long pl = idx / ((long)nw * nh);
        int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
        bilin_ac(oy, nh, H, y0, y1, ly0, ly1);
        bilin_ac(ox, nw, W, x0, x1, lx0, lx1);
        const float* s = x + pl * H * W;
        out[idx] = ly0 * (lx0 * s[(long)y0 * W + x0] + lx1 * s[(long)y0 * W + x1]) +
                   ly1 * (lx0 * s[(long)y1 * W + x0] + lx1 * s[(long)y1 * W + x1]);
    }
}
// adjoint of the resize: scatter-add the low-res gradient into the full-res gradient
__global__ void bilinear_ac_bwd_kernel(const float* __restrict__ g, long planes, int H, int W, int nh, int nw,
                                       float* __restrict__ gx) {
    long total = planes * nh * nw;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int ox = (int)(idx % nw);
        int oy = (int)((idx / nw) % nh);
        long pl = idx / ((long)nw * nh);
        int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
        bilin_ac(oy, nh, H, y0, y1, ly0, ly1);
        bilin_ac(ox, nw, W, x0, x1, lx0, lx1);
        float v = g[idx];
        float* s = gx + pl * H * W;
        atomicAdd(s + (long)y0 * W + x0, v * ly0 * lx0);
        atomicAdd(s + (long)y0 * W + x1, v * ly0 * lx1);
        atomicAdd(s + (long)y1 * W + x0, v * ly1 * lx0);
        atomicAdd(s + (long)y1 * W + x1, v * ly1 * lx1);
    }
}
// ------------------------------------------------------------------------------------------
// SSIM (pytorch_ssim/__init__.py:17-37): 11x11 Gaussian (sigma 1.5), zero 'same' padding,
// separable passes staged in shared memory; one 32x32 output tile of one (b,c) plane per CTA.
// ------------------------------------------------------------------------------------------
#define SSIM_T 32
#define SSIM_R 5
#define SSIM_IN (SSIM_T + 2 * SSIM_R)
__constant__ float c_gauss[11];
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                    int H, int W, double* __restrict__ out,
                                                    float* __restrict__ map) {
    __shared__ float sa[SSIM_IN][SSIM_IN + 1], sb[SSIM_IN][SSIM_IN + 1];
    __shared__ float hz[5][SSIM_IN][SSIM_T + 1];
    __shared__ double red[32];
    long pl = blockIdx.z;
    int y0 = blockIdx.y * SSIM_T, x0 = blockIdx.x * SSIM_T;
    const float* pa = a + pl * H * W;
    const float* pb = b + pl * H * W;
    for (int t = threadIdx.x; t < SSIM_IN * SSIM_IN; t += blockDim.x) {
        int r = t / SSIM_IN, c = t % SSIM_IN;
        int y = y0 + r - SSIM_R, x = x0 + c - SSIM_R;
        bool in = (y >= 0 && y < H && x >= 0 && x < W);
        sa[r][c] = in ? pa[(long)y * W + x] : 0.f;
        sb[r][c] = in ? pb[(long)y * W + x] : 0.f;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < SSIM_IN * SSIM_T; t += blockDim.x) {
        int r = t / SSIM_T, c = t % SSIM_T;
        float m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            float g = c_gauss[k], u = sa[r][c + k], v = sb[r][c + k];
            m1 += g * u; m2 += g * v; s11 += g * u * u; s22 += g * v * v; s12 += g * u * v;
        }
        hz[0][r][c] = m1; hz[1][r][c] = m2; hz[2][r][c] = s11; hz[3][r][c] = s22; hz[4][r][c] = s12;
    }
    __syncthreads();
    double acc = 0.0;
    for (int t = threadIdx.x; t < SSIM_T * SSIM_T; t += blockDim.x) {
        int r = t / SSIM_T, c = t % SSIM_T;
        int y = y0 + r, x = x0 + c;
        if (y >= H || x >= W) continue;
        float m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            float g = c_gauss[k];
            m1 += g * hz[0][r + k][c]; m2 += g * hz[1][r + k][c]; s11 += g * hz[2][r + k][c];
            s22 += g * hz[3][r + k][c]; s12 += g * hz[4][r + k][c];
        }
        float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
        float v1 = s11 - m11, v2 = s22 - m22, v12 = s12 - m12;
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        float s = ((2.f * m12 + C1) * (2.f * v12 + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2));
        if (map) map[pl * H * W + (long)y * W + x] = s;
        acc += (double)s;
    }
    acc = block_sum<double>(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}

// Rolling form: SSIM is bound by the fp32 pipe, not by HBM (five 11 x 11 separable windows = ~130 FMAs per pixel; at the HBM
// rate that would be 104 TFMA/s against 37), and the tiled kernel above spends its issue slots on shared-memory loads instead:
// 22 per pixel in the horizontal pass, 55 in the vertical one, plus 1.7x halo recomputation.  Here one thread owns ONE output
// column and walks down the rows of a chunk: the horizontal pass reads the current input row from a double-buffered
// shared-memory row (22 conflict-free loads, one barrier per row, the next row's global loads already in flight), the vertical
// pass is a SCATTER into 11 x 5 accumulators that live in registers - row r adds g[k] * hz(r) to the 11 output rows it
// touches, the oldest accumulator is complete and leaves.  The 11 phases of the accumulator ring are unrolled so that every
// register index is static.
#define SSIMR_NT 128
#define SSIMR_WW (32 + 2 * SSIM_R)                 // one warp's staged row: its 32 columns + the 5-pixel halo on both sides
__device__ __forceinline__ float2 ssim_f2(float x, float y) { return make_float2(x, y); }
__global__ void __launch_bounds__(SSIMR_NT, 4)
ssim_roll_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W, int rows, int chunks,
                 double* __restrict__ out, float* __restrict__ map) {
    // every WARP is on its own (32 output columns, its own double-buffered row in shared memory, __syncwarp only): a CTA-wide
    // row buffer cost one block barrier per row with four warps waiting on each other (measured r2q: 11.5 % of HBM)
    __shared__ float srow[SSIMR_NT / 32][2][2][SSIMR_WW + 2];
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pl = blockIdx.z, i0 = blockIdx.y * rows, i1 = min(i0 + rows, H), j0 = blockIdx.x * SSIMR_NT + warp * 32, j = j0 + lane;
    const float* pa = a + (long)pl * H * W;
    const float* pb = b + (long)pl * H * W;
    float g[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) g[k] = c_gauss[k];
    // columns this lane stages: s = lane (image column j0 - 5 + lane) and, for the first 10 lanes, s = 32 + lane
    const int c1 = j0 - SSIM_R + lane, c2 = c1 + 32;
    const bool in1 = c1 >= 0 && c1 < W, in2 = lane < 2 * SSIM_R && c2 < W;
    auto ld = [&](const float* p, int r, int c, bool in) { return (in && r >= 0 && r < H) ? __ldg(p + (long)r * W + c) : 0.f; };
    const int r0 = i0 - SSIM_R, r1 = i1 + SSIM_R;                 // input rows r0 .. r1 - 1 (zero outside the image)
    float (*buf)[2][SSIMR_WW + 2] = srow[warp];
    double total = 0.0;
    if (j0 < W) {                                                 // (warp-uniform: strips right of the image have nothing to do)
        float na1 = ld(pa, r0, c1, in1), nb1 = ld(pb, r0, c1, in1), na2 = ld(pa, r0, c2, in2), nb2 = ld(pb, r0, c2, in2);
        buf[r0 & 1][0][lane] = na1; buf[r0 & 1][1][lane] = nb1;
        if (lane < 2 * SSIM_R) { buf[r0 & 1][0][32 + lane] = na2; buf[r0 & 1][1][32 + lane] = nb2; }
        na1 = ld(pa, r0 + 1, c1, in1); nb1 = ld(pb, r0 + 1, c1, in1); na2 = ld(pa, r0 + 1, c2, in2); nb2 = ld(pb, r0 + 1, c2, in2);
        __syncwarp();
        // accumulator ring: 11 output rows in flight x (mu_a, mu_b | s_aa, s_bb | s_ab), the pairs as packed fp32 (FFMA2)
        float2 accm[11], accs[11];
        float accx[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) { accm[k] = ssim_f2(0.f, 0.f); accs[k] = ssim_f2(0.f, 0.f); accx[k] = 0.f; }
        float part = 0.f;
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        for (int rb = r0; rb < r1; rb += 11) {
#pragma unroll
            for (int ph = 0; ph < 11; ++ph) {
                const int r = rb + ph;
                if (r < r1) {                                        // warp-uniform
                    // stage row r + 1 (loaded during the previous iteration), start the loads of row r + 2
                    if (r + 1 < r1) {
                        const int nbuf = (r + 1) & 1;
                        buf[nbuf][0][lane] = na1; buf[nbuf][1][lane] = nb1;
                        if (lane < 2 * SSIM_R) { buf[nbuf][0][32 + lane] = na2; buf[nbuf][1][32 + lane] = nb2; }
                        na1 = ld(pa, r + 2, c1, in1); nb1 = ld(pb, r + 2, c1, in1); na2 = ld(pa, r + 2, c2, in2); nb2 = ld(pb, r + 2, c2, in2);
                    }
                    // horizontal pass of row r at this lane's column
                    const float* ua = buf[r & 1][0] + lane;
                    const float* ub = buf[r & 1][1] + lane;
                    float2 hm = ssim_f2(0.f, 0.f), hs = ssim_f2(0.f, 0.f);
                    float hx = 0.f;
#pragma unroll
                    for (int k = 0; k < 11; ++k) {
                        const float2 uv = ssim_f2(ua[k], ub[k]);
                        const float2 guv = __fmul2_rn(ssim_f2(g[k], g[k]), uv);
                        hm = __fadd2_rn(hm, guv);
                        hs = __ffma2_rn(guv, uv, hs);
                        hx = fmaf(guv.x, uv.y, hx);
                    }
                    // vertical scatter: input row r is tap k of output row r + 5 - k, which lives in ring slot (ph + 11 - k) % 11
                    // (slot (ph + 1) % 11 = tap 10 = the output row r - 5, complete after this update)
#pragma unroll
                    for (int k = 0; k < 11; ++k) {
                        const int sl = (ph + 11 - k) % 11;
                        const float2 gk = ssim_f2(g[k], g[k]);
                        accm[sl] = __ffma2_rn(gk, hm, accm[sl]);
                        accs[sl] = __ffma2_rn(gk, hs, accs[sl]);
                        accx[sl] = fmaf(g[k], hx, accx[sl]);
                    }
                    const int done = (ph + 1) % 11, y = r - SSIM_R;
                    if (y >= i0 && y < i1 && j < W) {
                        const float m1 = accm[done].x, m2 = accm[done].y;
                        const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
                        const float v1 = accs[done].x - m11, v2 = accs[done].y - m22, v12 = accx[done] - m12;
                        const float sv = ((2.f * m12 + C1) * (2.f * v12 + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2));
                        if (map) map[(long)pl * H * W + (long)y * W + j] = sv;
                        part += sv;
                    }
                    accm[done] = ssim_f2(0.f, 0.f); accs[done] = ssim_f2(0.f, 0.f); accx[done] = 0.f;
                    if (((r - r0) & 31) == 31) { total += (double)part; part = 0.f; }
                    __syncwarp();
                }
            }
        }
        total += (double)part;
    }
    total = block_sum<double>(total, red);
    if (tid == 0) atomicAdd(out, total);
}

// valid-depth mask of the Image Guidance step (models/I2D_model.py:223,226): out = (d < thr) ? 0 : 1
__global__ void below_mask_kernel(const float* __restrict__ d, long n, float thr, float* __restrict__ out) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = (d[i] < thr) ? 0.f : 1.f;
}

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
#define ST(s) ((cudaStream_t)(s))

extern "C" int dsr_below_mask(const float* depth, long n, float thr, float* out, void* stream) {
    DSR_REQUIRE(depth && out && n > 0, "bad arguments");
    below_mask_kernel<<<dsr_grid(n, TPB), TPB, 0, ST(stream)>>>(depth, n, thr, out);
    return dsr_check_launch("below_mask");
}
extern "C" int dsr_masked_sums(const float* d, const float* p, const float* m, long total, double* out3, void* stream) {
    DSR_REQUIRE(d && p && m && out3, "null pointer");
    masked_sums_kernel<<<dsr_grid(total, TPB), TPB, 0, ST(stream)>>>(d, p, m, total, out3);
    return dsr_check_launch("masked_sums");
}

extern "C" int dsr_bilinear_ac_fwd(const float* x, long planes, int H, int W, int nh, int nw, float* out, void* stream) {
    DSR_REQUIRE(x && out && nh > 0 && nw > 0, "bad arguments");
    bilinear_ac_fwd_kernel<<<dsr_grid(planes * nh * nw, TPB), TPB, 0, ST(stream)>>>(x, planes, H, W, nh, nw, out);
    return dsr_check_launch("bilinear_ac_fwd");
}
extern "C" int dsr_bilinear_ac_bwd(const float* g, long planes, int H, int W, int nh, int nw, float* gx, void* stream) {
    DSR_REQUIRE(g && gx, "null pointer");
    bilinear_ac_bwd_kernel<<<dsr_grid(planes * nh * nw, TPB), TPB, 0, ST(stream)>>>(g, planes, H, W, nh, nw, gx);
    return dsr_check_launch("bilinear_ac_bwd");
}
extern "C" int dsr_ssim_fwd(const float* a, const float* b, long planes, int H, int W, double* out_sum, float* map,
                            void* stream) {
    DSR_REQUIRE(a && b && out_sum, "null pointer");
    static bool init = false;
    if (!init) {
        float g[11], s = 0.f;
        for (int k = 0; k < 11; ++k) { g[k] = expf(-(float)((k - 5) * (k - 5)) / (2.f * 1.5f * 1.5f)); s += g[k]; }
        for (int k = 0; k < 11; ++k) g[k] /= s;
        if (cudaMemcpyToSymbol(c_gauss, g, sizeof(g)) != cudaSuccess) { dsr_set_error("ssim: constant upload failed"); return DSR_ERR_CUDA; }
        init = true;
    }
    static int mode = -1;                          // DSR_SSIM_KERNEL=0: the first-generation tiled kernel (A/B runs)
    if (mode < 0) { const char* e = getenv("DSR_SSIM_KERNEL"); mode = e ? atoi(e) : 1; }
    if (mode != 0 && planes <= 65535) {
        // chunks of rows: enough CTAs for ~2 waves of 4 CTAs per SM, at least 32 rows each (10 halo rows per chunk)
        const long strips = (long)dsr_cdiv(W, SSIMR_NT) * planes, want = (long)dsr_num_sms() * 8;
        long chunks = (want + strips - 1) / strips;
        if (chunks < 1) chunks = 1;
        long rows = (H + chunks - 1) / chunks;
        if (rows < 32) rows = 32;
        if (rows > H) rows = H;
        chunks = (H + rows - 1) / rows;
        if (chunks <= 65535) {
            dim3 grid(dsr_cdiv(W, SSIMR_NT), (unsigned)chunks, (unsigned)planes);
            ssim_roll_kernel<<<grid, SSIMR_NT, 0, ST(stream)>>>(a, b, H, W, (int)rows, (int)chunks, out_sum, map);
            return dsr_check_launch("ssim_fwd");
        }
    }
    dim3 grid(dsr_cdiv(W, SSIM_T), dsr_cdiv(H, SSIM_T), (unsigned)planes);
    ssim_kernel<<<grid, 256, 0, ST(stream)>>>(a, b, H, W, out_sum, map);
    return dsr_check_launch("ssim_fwd");
}
