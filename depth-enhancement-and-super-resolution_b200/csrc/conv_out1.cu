// Single-output-channel convolutions (the depth heads): Conv2d k x k stride 1 -> 1 channel (translation_network.py:495,
// G_A_d's 7x7 64 -> 1 + tanh) and ConvTranspose2d 4x4 stride 2 pad 1 -> 1 channel (networks.py:553, the U-Net heads
// 128 -> 1 + tanh).  As GEMMs these have N = 1 (0.2 .. 0.4 GMAC per call, 500 / 200 us on the tensor path with a 16-wide
// MMA of which one column is real); they are bandwidth-bound reductions over the input, so they run on CUDA cores:
// fp32 NHWC input read once into shared memory (with the fused norm-apply + activation prologue and the padding mode
// applied on the way in), channel-quad-major tiles so a warp's float4 reads are conflict-free, weights broadcast from
// shared memory, bias + tanh in the epilogue.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/dsr_b200.h"

#define ST(s) ((cudaStream_t)(s))
#define O1_TILE 16
#define O1_CC 16                 // channels per shared-memory chunk

__device__ __forceinline__ int o1_pad_src(int i, int n, int mode) {       // index in [-(pad), n + pad) -> source or -1
    if (i >= 0 && i < n) return i;
    if (mode == DSR_PAD_ZERO) return -1;
    if (mode == DSR_PAD_REFLECT) { i = i < 0 ? -i : 2 * (n - 1) - i; return (i >= 0 && i < n) ? i : -1; }
    return i < 0 ? 0 : n - 1;
}

__device__ __forceinline__ float o1_prologue(float v, const float* __restrict__ prm, long NC, long k, int act, float slope) {
    if (prm) v = (v - prm[k]) * prm[NC + k] + prm[2 * NC + k];
    if (act == DSR_ACT_RELU) v = v > 0.f ? v : 0.f;
    else if (act == DSR_ACT_LRELU) v = v > 0.f ? v : slope * v;
    return v;
}

// 4 consecutive channels of one source pixel with the fused norm-apply + activation prologue: 16-byte loads when the
// channel count allows it (the per-element form cost 4 + 12 scalar loads)
__device__ __forceinline__ float4 o1_load4(const float* __restrict__ src, int c, int C, const float* __restrict__ prm, long NC,
                                           long k, int act, float slope, bool vec) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec) {
        const float4 a = ld4(src);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        if (prm) {
            const float4 m = ld4(prm + k), sc = ld4(prm + NC + k), sh = ld4(prm + 2 * NC + k);
            v[0] = (v[0] - m.x) * sc.x + sh.x; v[1] = (v[1] - m.y) * sc.y + sh.y;
            v[2] = (v[2] - m.z) * sc.z + sh.z; v[3] = (v[3] - m.w) * sc.w + sh.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (act == DSR_ACT_RELU) v[e] = v[e] > 0.f ? v[e] : 0.f;
            else if (act == DSR_ACT_LRELU) v[e] = v[e] > 0.f ? v[e] : slope * v[e];
        }
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (c + e < C) v[e] = o1_prologue(src[e], prm, NC, k + e, act, slope);
    }
    return make_float4(v[0], v[1], v[2], v[3]);
}

// stride-1 R x S convolution to one channel.  Block = 16 x 16 output pixels; per 16-channel chunk the (16+R-1)^2 input
// patch sits in shared memory as [channel quad][y][x] float4.
template <int K>        // K = R = S when known at compile time (7: constant divisors in the staging loop), 0 = generic
__global__ void __launch_bounds__(256)
conv_out1_s1_kernel(const float* __restrict__ x, int N, int H, int W, int C, const float* __restrict__ prm, int act_in,
                    float slope, const float* __restrict__ w /* [C][R][S] */, const float* __restrict__ bias, int R_, int S_,
                    int pad, int pad_mode, int act_out, float* __restrict__ out, int Ho, int Wo) {
    extern __shared__ float4 o1_sm[];
    const int R = K ? K : R_, S = K ? K : S_;
    const int PH = O1_TILE + R - 1, PW = O1_TILE + S - 1;
    const bool vec = (C & 3) == 0 && !((uintptr_t)x & 15) && (!prm || !((uintptr_t)prm & 15));
    float4* tile = o1_sm;                                   // [4][PH][PW]
    float4* wsm = o1_sm + 4 * PH * PW;                      // [R*S][4]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int tiles_w = (Wo + O1_TILE - 1) / O1_TILE, tiles_h = (Ho + O1_TILE - 1) / O1_TILE;
    const long NC = (long)N * C;
    for (int t = blockIdx.x; t < N * tiles_h * tiles_w; t += gridDim.x) {
        const int n = t / (tiles_h * tiles_w), r0 = t - n * tiles_h * tiles_w;
        const int h0 = (r0 / tiles_w) * O1_TILE, w0 = (r0 % tiles_w) * O1_TILE;
        float acc = 0.f;
        for (int c0 = 0; c0 < C; c0 += O1_CC) {
            __syncthreads();
            for (int i = tid; i < 4 * PH * PW; i += 256) {
                const int cq = i / (PH * PW), pq = i - cq * PH * PW;
                const int py = pq / PW, px = pq - py * PW;
                const int sy = o1_pad_src(h0 + py - pad, H, pad_mode), sx = o1_pad_src(w0 + px - pad, W, pad_mode);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sy >= 0 && sx >= 0) {
                    const int c = c0 + cq * 4;
                    if (c < C) v = o1_load4(x + ((long)(n * H + sy) * W + sx) * C + c, c, C, prm, NC, (long)n * C + c, act_in, slope, vec);
                }
                tile[i] = v;
            }
            for (int i = tid; i < R * S * 4; i += 256) {
                const int tap = i >> 2, cq = i & 3, c = c0 + cq * 4;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (c + e < C) ? w[(long)(c + e) * R * S + tap] : 0.f;
                wsm[i] = make_float4(v[0], v[1], v[2], v[3]);
            }
            __syncthreads();
            for (int r = 0; r < R; ++r)
                for (int s = 0; s < S; ++s) {
                    const int tap = r * S + s;
#pragma unroll
                    for (int cq = 0; cq < 4; ++cq) {
                        const float4 a = tile[(cq * PH + ty + r) * PW + tx + s], b = wsm[tap * 4 + cq];
                        acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
                    }
                }
        }
        const int h = h0 + ty, ww = w0 + tx;
        if (h < Ho && ww < Wo) {
            float o = acc + (bias ? bias[0] : 0.f);
            if (act_out == DSR_ACT_TANH) o = tanhf(o);
            out[((long)n * Ho + h) * Wo + ww] = o;
        }
    }
}

// The same convolution register-blocked for the K x K heads the step runs (K = 7: G_A_d's 64 -> 1 head and the data gradients
// of the one-channel 7x7 first layers).  conv_out1_s1_kernel issues 8 LDS.128 per 16 FMA instructions (one output pixel per
// thread: every tap re-reads its activations and its weights) and runs at ~12 % of the fp32 pipe, bound by shared-memory
// wavefronts.  Here a thread owns O1R_RB = 4 vertically adjacent output pixels: per (channel quad, kernel column s) it loads
// the RB + K - 1 = 10 activations of its column ONCE into registers and reuses them across the K kernel rows and the RB
// outputs - 10 conflict-free LDS.128 (a half-warp reads 256 contiguous bytes) + K broadcast weight loads per K * RB * 4 = 112
// FMAs, i.e. 47 shared-memory wavefronts per 28 FMA issue cycles instead of 32 per 4.  Block = 16 (w) x 64 (h) outputs, 8
// channels per shared-memory chunk (50 KB: four blocks per SM); lanes stage (pixel, quad) pairs so that a pixel's 32 bytes of
// a chunk are one sector.
#define O1R_RB 4
#define O1R_CQ 2
template <int K>
__global__ void __launch_bounds__(256)
conv_out1_s1_rb_kernel(const float* __restrict__ x, int N, int H, int W, int C, const float* __restrict__ prm, int act_in,
                       float slope, const float* __restrict__ w /* [C][K][K] */, const float* __restrict__ bias, int pad,
                       int pad_mode, int act_out, float* __restrict__ out, int Ho, int Wo) {
    extern __shared__ float4 o1_sm[];
    constexpr int TH = 16 * O1R_RB, PH = TH + K - 1, PW = O1_TILE + K - 1, PP = PH * PW;
    const bool vec = (C & 3) == 0 && !((uintptr_t)x & 15) && (!prm || !((uintptr_t)prm & 15));
    float4* tile = o1_sm;                                   // [CQ][PH][PW]
    float4* wsm = o1_sm + O1R_CQ * PP;                      // [K*K][CQ]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int tiles_w = (Wo + O1_TILE - 1) / O1_TILE, tiles_h = (Ho + TH - 1) / TH;
    const long NC = (long)N * C;
    for (int t = blockIdx.x; t < N * tiles_h * tiles_w; t += gridDim.x) {
        const int n = t / (tiles_h * tiles_w), r0 = t - n * tiles_h * tiles_w;
        const int h0 = (r0 / tiles_w) * TH, w0 = (r0 % tiles_w) * O1_TILE;
        float acc[O1R_RB];
#pragma unroll
        for (int i = 0; i < O1R_RB; ++i) acc[i] = 0.f;
        for (int c0 = 0; c0 < C; c0 += 4 * O1R_CQ) {
            __syncthreads();
            for (int i = tid; i < O1R_CQ * PP; i += 256) {
                const int pq = i / O1R_CQ, cq = i - pq * O1R_CQ;
                const int py = pq / PW, px = pq - py * PW;
                const int sy = o1_pad_src(h0 + py - pad, H, pad_mode), sx = o1_pad_src(w0 + px - pad, W, pad_mode);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sy >= 0 && sx >= 0) {
                    const int c = c0 + cq * 4;
                    if (c < C) v = o1_load4(x + ((long)(n * H + sy) * W + sx) * C + c, c, C, prm, NC, (long)n * C + c, act_in, slope, vec);
                }
                tile[cq * PP + pq] = v;
            }
            for (int i = tid; i < K * K * O1R_CQ; i += 256) {
                const int tap = i / O1R_CQ, cq = i - tap * O1R_CQ, c = c0 + cq * 4;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (c + e < C) ? w[(long)(c + e) * K * K + tap] : 0.f;
                wsm[i] = make_float4(v[0], v[1], v[2], v[3]);
            }
            __syncthreads();
#pragma unroll
            for (int cq = 0; cq < O1R_CQ; ++cq) {
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    float4 a[O1R_RB + K - 1];
#pragma unroll
                    for (int j = 0; j < O1R_RB + K - 1; ++j) a[j] = tile[cq * PP + (ty * O1R_RB + j) * PW + tx + s];
#pragma unroll
                    for (int r = 0; r < K; ++r) {
                        const float4 b = wsm[(r * K + s) * O1R_CQ + cq];
#pragma unroll
                        for (int i = 0; i < O1R_RB; ++i)
                            acc[i] += a[i + r].x * b.x + a[i + r].y * b.y + a[i + r].z * b.z + a[i + r].w * b.w;
                    }
                }
            }
        }
        const int ww = w0 + tx;
        if (ww < Wo) {
            const float bv = bias ? bias[0] : 0.f;
#pragma unroll
            for (int i = 0; i < O1R_RB; ++i) {
                const int h = h0 + ty * O1R_RB + i;
                if (h < Ho) {
                    float o = acc[i] + bv;
                    if (act_out == DSR_ACT_TANH) o = tanhf(o);
                    out[((long)n * Ho + h) * Wo + ww] = o;
                }
            }
        }
    }
}

// ConvTranspose2d 4x4 stride 2 pad 1 to one channel: each thread owns one input pixel position (h, w) and produces its four
// output phases out[2h+a][2w+b] = sum_{dr,ds in {0,1}} sum_c x[h-1+a+dr][w-1+b+ds][c] * W[c][3-a-2dr][3-b-2ds].
__global__ void __launch_bounds__(256)
convT4_out1_kernel(const float* __restrict__ x, int N, int H, int W, int C, const float* __restrict__ prm, int act_in,
                   float slope, const float* __restrict__ w /* [C][4][4] */, const float* __restrict__ bias, int act_out,
                   float* __restrict__ out /* N x 2H x 2W */) {
    extern __shared__ float4 o1_sm[];
    constexpr int PH = O1_TILE + 2, PW = O1_TILE + 2;
    float4* tile = o1_sm;                                   // [4][PH][PW]
    float4* wsm = o1_sm + 4 * PH * PW;                      // [16][4]
    const bool vec = (C & 3) == 0 && !((uintptr_t)x & 15) && (!prm || !((uintptr_t)prm & 15));
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int tiles_w = (W + O1_TILE - 1) / O1_TILE, tiles_h = (H + O1_TILE - 1) / O1_TILE;
    const long NC = (long)N * C;
    for (int t = blockIdx.x; t < N * tiles_h * tiles_w; t += gridDim.x) {
        const int n = t / (tiles_h * tiles_w), r0 = t - n * tiles_h * tiles_w;
        const int h0 = (r0 / tiles_w) * O1_TILE, w0 = (r0 % tiles_w) * O1_TILE;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c0 = 0; c0 < C; c0 += O1_CC) {
            __syncthreads();
            for (int i = tid; i < 4 * PH * PW; i += 256) {
                const int cq = i / (PH * PW), pq = i - cq * PH * PW;
                const int py = pq / PW, px = pq - py * PW;
                const int sy = h0 + py - 1, sx = w0 + px - 1;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
                    const int c = c0 + cq * 4;
                    if (c < C) v = o1_load4(x + ((long)(n * H + sy) * W + sx) * C + c, c, C, prm, NC, (long)n * C + c, act_in, slope, vec);
                }
                tile[i] = v;
            }
            for (int i = tid; i < 64; i += 256) {
                const int tap = i >> 2, cq = i & 3, c = c0 + cq * 4;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (c + e < C) ? w[(long)(c + e) * 16 + tap] : 0.f;
                wsm[i] = make_float4(v[0], v[1], v[2], v[3]);
            }
            __syncthreads();
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int dr = 0; dr < 2; ++dr)
#pragma unroll
                        for (int ds = 0; ds < 2; ++ds) {
                            const int tap = (3 - a - 2 * dr) * 4 + (3 - b - 2 * ds);
#pragma unroll
                            for (int cq = 0; cq < 4; ++cq) {
                                const float4 xv = tile[(cq * PH + ty + a + dr) * PW + tx + b + ds], wv = wsm[tap * 4 + cq];
                                acc[a * 2 + b] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
                            }
                        }
        }
        const int h = h0 + ty, ww = w0 + tx;
        if (h < H && ww < W) {
            const float bv = bias ? bias[0] : 0.f;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    float o = acc[a * 2 + b] + bv;
                    if (act_out == DSR_ACT_TANH) o = tanhf(o);
                    out[((long)n * 2 * H + 2 * h + a) * 2 * W + 2 * ww + b] = o;
                }
        }
    }
}

// The transposed head register-blocked the same way: a thread owns O1R_RB = 4 vertically adjacent input positions (16 outputs),
// per channel quad it loads its 6 x 3 activations once (18 conflict-free LDS.128) + the 16 taps (broadcast) for 256 FMAs
// (convT4_out1_kernel: 32 LDS.128 per 64 FMAs).  Block = 128 threads = 16 (w) x 32 (h) input positions, 8 channels per chunk.
__global__ void __launch_bounds__(128)
convT4_out1_rb_kernel(const float* __restrict__ x, int N, int H, int W, int C, const float* __restrict__ prm, int act_in,
                      float slope, const float* __restrict__ w /* [C][4][4] */, const float* __restrict__ bias, int act_out,
                      float* __restrict__ out /* N x 2H x 2W */) {
    extern __shared__ float4 o1_sm[];
    constexpr int TH = 8 * O1R_RB, PH = TH + 2, PW = O1_TILE + 2, PP = PH * PW;
    float4* tile = o1_sm;                                   // [CQ][PH][PW]
    float4* wsm = o1_sm + O1R_CQ * PP;                      // [16][CQ]
    const bool vec = (C & 3) == 0 && !((uintptr_t)x & 15) && (!prm || !((uintptr_t)prm & 15));
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int tiles_w = (W + O1_TILE - 1) / O1_TILE, tiles_h = (H + TH - 1) / TH;
    const long NC = (long)N * C;
    for (int t = blockIdx.x; t < N * tiles_h * tiles_w; t += gridDim.x) {
        const int n = t / (tiles_h * tiles_w), r0 = t - n * tiles_h * tiles_w;
        const int h0 = (r0 / tiles_w) * TH, w0 = (r0 % tiles_w) * O1_TILE;
        float acc[O1R_RB][4];
#pragma unroll
        for (int i = 0; i < O1R_RB; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        for (int c0 = 0; c0 < C; c0 += 4 * O1R_CQ) {
            __syncthreads();
            for (int i = tid; i < O1R_CQ * PP; i += 128) {
                const int pq = i / O1R_CQ, cq = i - pq * O1R_CQ;
                const int py = pq / PW, px = pq - py * PW;
                const int sy = h0 + py - 1, sx = w0 + px - 1;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
                    const int c = c0 + cq * 4;
                    if (c < C) v = o1_load4(x + ((long)(n * H + sy) * W + sx) * C + c, c, C, prm, NC, (long)n * C + c, act_in, slope, vec);
                }
                tile[cq * PP + pq] = v;
            }
            for (int i = tid; i < 16 * O1R_CQ; i += 128) {
                const int tap = i / O1R_CQ, cq = i - tap * O1R_CQ, c = c0 + cq * 4;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (c + e < C) ? w[(long)(c + e) * 16 + tap] : 0.f;
                wsm[i] = make_float4(v[0], v[1], v[2], v[3]);
            }
            __syncthreads();
#pragma unroll
            for (int cq = 0; cq < O1R_CQ; ++cq) {
                float4 a[O1R_RB + 2][3];
#pragma unroll
                for (int j = 0; j < O1R_RB + 2; ++j)
#pragma unroll
                    for (int k = 0; k < 3; ++k) a[j][k] = tile[cq * PP + (ty * O1R_RB + j) * PW + tx + k];
#pragma unroll
                for (int pa = 0; pa < 2; ++pa)
#pragma unroll
                    for (int pb = 0; pb < 2; ++pb)
#pragma unroll
                        for (int dr = 0; dr < 2; ++dr)
#pragma unroll
                            for (int ds = 0; ds < 2; ++ds) {
                                const float4 wv = wsm[((3 - pa - 2 * dr) * 4 + (3 - pb - 2 * ds)) * O1R_CQ + cq];
#pragma unroll
                                for (int i = 0; i < O1R_RB; ++i) {
                                    const float4 xv = a[i + pa + dr][pb + ds];
                                    acc[i][pa * 2 + pb] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
                                }
                            }
            }
        }
        const int ww = w0 + tx;
        if (ww < W) {
            const float bv = bias ? bias[0] : 0.f;
#pragma unroll
            for (int i = 0; i < O1R_RB; ++i) {
                const int h = h0 + ty * O1R_RB + i;
                if (h < H) {
#pragma unroll
                    for (int pa = 0; pa < 2; ++pa) {
                        float o0 = acc[i][pa * 2] + bv, o1 = acc[i][pa * 2 + 1] + bv;
                        if (act_out == DSR_ACT_TANH) { o0 = tanhf(o0); o1 = tanhf(o1); }
                        *reinterpret_cast<float2*>(out + ((long)n * 2 * H + 2 * h + pa) * 2 * W + 2 * ww) = make_float2(o0, o1);
                    }
                }
            }
        }
    }
}

extern "C" int dsr_conv_out1(const float* x, int N, int H, int W, int C, const float* prm, int act_in, float slope,
                             const float* w, const float* bias, int R, int S, int pad, int pad_mode, int transposed,
                             int act_out, float* out, void* stream) {
    DSR_REQUIRE(x && w && out && N > 0 && H > 0 && W > 0 && C > 0, "bad arguments");
    DSR_REQUIRE((long)N * H * W < (1L << 31) / 4, "tensor too large for 32-bit pixel indices");
    if (transposed) {
        DSR_REQUIRE(R == 4 && S == 4 && pad == 1, "transposed variant: 4x4, stride 2, padding 1");
        {
            const char* e = getenv("DSR_OUT1_RB");
            if (H >= 8 * O1R_RB && !((uintptr_t)out & 7) && (!e || atoi(e) != 0)) {
                const int tiles_rb = N * dsr_cdiv(H, 8 * O1R_RB) * dsr_cdiv(W, O1_TILE);
                const size_t smem_rb = (size_t)(O1R_CQ * (8 * O1R_RB + 2) * (O1_TILE + 2) + 16 * O1R_CQ) * sizeof(float4);
                const int cap_rb = dsr_num_sms() * 8;
                convT4_out1_rb_kernel<<<tiles_rb < cap_rb ? tiles_rb : cap_rb, 128, smem_rb, ST(stream)>>>(x, N, H, W, C, prm, act_in, slope, w,
                                                                                                       bias, act_out, out);
                return dsr_check_launch("conv_out1 (transposed, register-blocked)");
            }
        }
        const int tiles = N * dsr_cdiv(H, O1_TILE) * dsr_cdiv(W, O1_TILE);
        const size_t smem = (4 * (O1_TILE + 2) * (O1_TILE + 2) + 64) * sizeof(float4);
        const int cap = dsr_num_sms() * 4;
        convT4_out1_kernel<<<tiles < cap ? tiles : cap, 256, smem, ST(stream)>>>(x, N, H, W, C, prm, act_in, slope, w, bias, act_out, out);
        return dsr_check_launch("conv_out1 (transposed)");
    }
    DSR_REQUIRE(R >= 1 && S >= 1 && R <= 9 && S <= 9 && pad >= 0, "kernel size 1..9");
    DSR_REQUIRE(pad_mode != DSR_PAD_REFLECT || (pad < H && pad < W), "reflect padding needs pad < size");
    const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
    const int tiles = N * dsr_cdiv(Ho, O1_TILE) * dsr_cdiv(Wo, O1_TILE);
    const size_t smem = (4 * (O1_TILE + R - 1) * (O1_TILE + S - 1) + R * S * 4) * sizeof(float4);
    const int cap = dsr_num_sms() * 4;
    {
        const char* e = getenv("DSR_OUT1_RB");           // 0: the one-pixel-per-thread kernel (A/B)
        if (R == 7 && S == 7 && Ho >= 16 * O1R_RB && (!e || atoi(e) != 0)) {
            constexpr int PHr = 16 * O1R_RB + 6, PWr = O1_TILE + 6;
            const size_t smem_rb = (size_t)(O1R_CQ * PHr * PWr + 49 * O1R_CQ) * sizeof(float4);
            static bool attr_set = false;
            if (!attr_set) {
                if (cudaFuncSetAttribute(conv_out1_s1_rb_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rb) != cudaSuccess) {
                    dsr_set_error("conv_out1: cannot raise the shared-memory limit"); return DSR_ERR_CUDA;
                }
                attr_set = true;
            }
            const int tiles_rb = N * dsr_cdiv(Ho, 16 * O1R_RB) * dsr_cdiv(Wo, O1_TILE);
            conv_out1_s1_rb_kernel<7><<<tiles_rb < cap ? tiles_rb : cap, 256, smem_rb, ST(stream)>>>(x, N, H, W, C, prm, act_in, slope, w, bias,
                                                                                                 pad, pad_mode, act_out, out, Ho, Wo);
            return dsr_check_launch("conv_out1 (register-blocked)");
        }
    }
    if (R == 7 && S == 7)
        conv_out1_s1_kernel<7><<<tiles < cap ? tiles : cap, 256, smem, ST(stream)>>>(x, N, H, W, C, prm, act_in, slope, w, bias, R, S, pad,
                                                                                  pad_mode, act_out, out, Ho, Wo);
    else
        conv_out1_s1_kernel<0><<<tiles < cap ? tiles : cap, 256, smem, ST(stream)>>>(x, N, H, W, C, prm, act_in, slope, w, bias, R, S, pad,
                                                                                  pad_mode, act_out, out, Ho, Wo);
    return dsr_check_launch("conv_out1");
}

// Data gradient of a stride-1 convolution that has ONE output channel (the generator heads 64 -> 1 of the translation block,
// the last discriminator layer 512 -> 1): gx[n][i][j][c] = sum_{r,s} g[n][i + off - r][j + off - s] * W[c][r][s]  (g = 0 outside).
// An outer product per tap, bound by writing gx: one thread = one pixel, 16 channels at a time in registers, the gradient
// tile and the [tap][channel] weights in shared memory.  (The CUDA-core fallback spent ~5 ms per 7x7 head at 6 x 256 x 256.)
__global__ void __launch_bounds__(256)
conv1_dgrad_kernel(const float* __restrict__ g, int N, int Ho, int Wo, const float* __restrict__ w /* [C][R][S] */, int C, int R,
                   int S, int off, float* __restrict__ gx, int Hx, int Wx) {
    extern __shared__ float d1_sm[];
    const int PH = O1_TILE + R - 1, PW = O1_TILE + S - 1;
    float* tile = d1_sm;                                    // [PH][PW] gradient patch
    float* wsm = d1_sm + ((PH * PW + 3) & ~3);              // [R*S][C]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int i = tid; i < R * S * C; i += 256) {
        const int tap = i / C, c = i - tap * C;
        wsm[i] = w[(long)c * R * S + tap];
    }
    const int tiles_w = (Wx + O1_TILE - 1) / O1_TILE, tiles_h = (Hx + O1_TILE - 1) / O1_TILE;
    for (int t = blockIdx.x; t < N * tiles_h * tiles_w; t += gridDim.x) {
        const int n = t / (tiles_h * tiles_w), r0 = t - n * tiles_h * tiles_w;
        const int i0 = (r0 / tiles_w) * O1_TILE, j0 = (r0 % tiles_w) * O1_TILE;
        __syncthreads();
        for (int q = tid; q < PH * PW; q += 256) {
            const int py = q / PW, px = q - py * PW;
            const int gy = i0 + off - (R - 1) + py, gxx = j0 + off - (S - 1) + px;
            tile[q] = (gy >= 0 && gy < Ho && gxx >= 0 && gxx < Wo) ? g[((long)n * Ho + gy) * Wo + gxx] : 0.f;
        }
        __syncthreads();
        const int i = i0 + ty, j = j0 + tx;
        if (i >= Hx || j >= Wx) continue;
        float* o = gx + (((long)n * Hx + i) * Wx + j) * C;
        for (int c0 = 0; c0 < C; c0 += 16) {
            float acc[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = 0.f;
            for (int r = 0; r < R; ++r)
                for (int s = 0; s < S; ++s) {
                    const float gv = tile[(ty + (R - 1) - r) * PW + tx + (S - 1) - s];      // g[i + off - r][j + off - s]
                    const float* wr = wsm + (r * S + s) * C + c0;
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const float4 wv = *reinterpret_cast<const float4*>(wr + e);
                        acc[e] += gv * wv.x; acc[e + 1] += gv * wv.y; acc[e + 2] += gv * wv.z; acc[e + 3] += gv * wv.w;
                    }
                }
#pragma unroll
            for (int e = 0; e < 16; e += 4) st4(o + c0 + e, make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]));
        }
    }
}
extern "C" int dsr_conv1_dgrad(const float* g, int N, int Ho, int Wo, const float* w, int C, int R, int S, int off, float* gx,
                               int Hx, int Wx, void* stream) {
    DSR_REQUIRE(g && w && gx && N > 0 && Ho > 0 && Wo > 0 && Hx > 0 && Wx > 0, "bad arguments");
    DSR_REQUIRE(C >= 16 && (C & 15) == 0 && R >= 1 && S >= 1 && R <= 9 && S <= 9 && !((uintptr_t)gx & 15), "C % 16 == 0, kernel size 1..9");
    const size_t smem = ((((size_t)(O1_TILE + R - 1) * (O1_TILE + S - 1) + 3) & ~(size_t)3) + (size_t)R * S * C) * sizeof(float);
    DSR_REQUIRE(smem <= 200 * 1024, "weights do not fit shared memory");
    static size_t raised = 0;
    if (smem > 48 * 1024 && smem > raised) {
        if (cudaFuncSetAttribute(conv1_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            dsr_set_error("conv1_dgrad: cannot raise dynamic shared memory to %d", (int)smem);
            return DSR_ERR_CUDA;
        }
        raised = smem;
    }
    const int tiles = N * dsr_cdiv(Hx, O1_TILE) * dsr_cdiv(Wx, O1_TILE);
    const int cap = dsr_num_sms() * 4;
    conv1_dgrad_kernel<<<tiles < cap ? tiles : cap, 256, smem, ST(stream)>>>(g, N, Ho, Wo, w, C, R, S, off, gx, Hx, Wx);
    return dsr_check_launch("conv1_dgrad");
}

// Border term of the data gradient of a stride-2, padding-1 Conv2d whose padding is replicate / reflect (the generator
// encoders, translation_network.py:478).  With gxp = the gradient w.r.t. the explicitly padded input (H+2 x W+2),
//   gx = gxp[1..H][1..W]  (the zero-padding data gradient: the tcgen05 transposed-phase GEMM)
//      + the one-pixel FRAME of gxp folded onto the source pixels the padding mode read it from.
// The frame is 2(W+2)+2H pixels per image with <= 2x2 taps each, so this kernel evaluates it directly: one warp per
// affected source pixel, lanes over input channels, frame pixels summed in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
conv_s2_border_dgrad_kernel(const float* __restrict__ g, int N, int Ho, int Wo, int Co, const float* __restrict__ w /* [Co][Ci][R][S] */,
                            int Ci, int R, int S, int mode, float* __restrict__ gx, int H, int W) {
    const int lo_h = mode == DSR_PAD_REFLECT ? 1 : 0, hi_h = mode == DSR_PAD_REFLECT ? H - 2 : H - 1;
    const int lo_w = mode == DSR_PAD_REFLECT ? 1 : 0, hi_w = mode == DSR_PAD_REFLECT ? W - 2 : W - 1;
    const int nr = lo_h == hi_h ? 1 : 2, nc = lo_w == hi_w ? 1 : 2;
    const int per_img = nr * W + (H - nr) * nc;
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long item = warp; item < (long)N * per_img; item += nwarps) {
        const int n = (int)(item / per_img), q = (int)(item - (long)n * per_img);
        int i, j;
        if (q < nr * W) { i = q / W == 0 ? lo_h : hi_h; j = q - (q / W) * W; }
        else {
            const int q2 = q - nr * W, k = q2 / nc;
            i = k; if (i >= lo_h) ++i; if (nr == 2 && i >= hi_h) ++i;
            j = (q2 - k * nc) == 0 ? lo_w : hi_w;
        }
        // padded rows / columns (frame or centre) that read source row i / column j
        int ph[3], pw[3], nph = 0, npw = 0;
        ph[nph++] = i + 1; if (i == lo_h) ph[nph++] = 0; if (i == hi_h) ph[nph++] = H + 1;
        pw[npw++] = j + 1; if (j == lo_w) pw[npw++] = 0; if (j == hi_w) pw[npw++] = W + 1;
        for (int c = lane; c < Ci; c += 32) {
            float acc = 0.f;
            for (int a = 0; a < nph; ++a)
                for (int b = 0; b < npw; ++b) {
                    if (a == 0 && b == 0) continue;                       // the centre pixel: already in gx
                    for (int r = ph[a] & 1; r < R; r += 2) {
                        const int oh = (ph[a] - r) >> 1;
                        if (ph[a] - r < 0 || oh >= Ho) continue;
                        for (int s = pw[b] & 1; s < S; s += 2) {
                            const int ow = (pw[b] - s) >> 1;
                            if (pw[b] - s < 0 || ow >= Wo) continue;
                            const float* gp = g + (((long)n * Ho + oh) * Wo + ow) * Co;
                            const float* wp = w + ((long)c * R + r) * S + s;
                            for (int co = 0; co < Co; ++co) acc = fmaf(gp[co], wp[(long)co * Ci * R * S], acc);
                        }
                    }
                }
            gx[(((long)n * H + i) * W + j) * Ci + c] += acc;
        }
    }
}
extern "C" int dsr_conv_s2_border_dgrad(const float* g, int N, int Ho, int Wo, int Co, const float* w, int Ci, int R, int S,
                                        int pad_mode, float* gx, int H, int W, void* stream) {
    DSR_REQUIRE(g && w && gx && N > 0 && Ho > 0 && Wo > 0 && Co > 0 && Ci > 0 && H > 0 && W > 0, "bad arguments");
    DSR_REQUIRE(R >= 1 && S >= 1 && Ho == (H + 2 - R) / 2 + 1 && Wo == (W + 2 - S) / 2 + 1, "stride 2, padding 1 geometry");
    DSR_REQUIRE(pad_mode == DSR_PAD_REPLICATE || (pad_mode == DSR_PAD_REFLECT && H >= 2 && W >= 2), "replicate or reflect padding");
    const long warps = (long)N * (2L * W + 2L * H);
    conv_s2_border_dgrad_kernel<<<dsr_grid(warps * 32, 256), 256, 0, ST(stream)>>>(g, N, Ho, Wo, Co, w, Ci, R, S, pad_mode, gx, H, W);
    return dsr_check_launch("conv_s2_border_dgrad");
}
