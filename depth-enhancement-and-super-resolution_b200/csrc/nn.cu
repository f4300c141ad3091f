// Layout, padding, activation, normalisation and optimizer kernels (HBM-bound, NHWC fp32).
// Reference call sites: models/networks.py:30 (InstanceNorm2d affine=False), :378-415 (ReflectionPad2d,
// ReLU, Tanh), :546-557 (LeakyReLU 0.2 / ReLU), :629 (skip concat); models/translation_network.py:46
// (GroupNorm(8, C, affine=True)), :472-478 (replicate padding); models/main_model.py:176 (Adam).
#include "common.cuh"
#include "../../include/dsr_b200.h"

#define TPB 256
#define ST(s) ((cudaStream_t)(s))

// ------------------------------------------------------------------------------------------
// NCHW <-> NHWC  (x: [N][C][P]  <->  y: [N][P][C]); 32x32 smem tile transpose
// ------------------------------------------------------------------------------------------
__global__ void transpose_cp_kernel(const float* __restrict__ x, float* __restrict__ y, int C, long P, int to_nhwc) {
    __shared__ float tile[32][33];
    int n = blockIdx.z;
    long p0 = (long)blockIdx.x * 32;
    int c0 = blockIdx.y * 32;
    const float* xs = x + (long)n * C * P;
    float* ys = y + (long)n * C * P;
    if (to_nhwc) {  // read [c][p] coalesced along p, write [p][c] coalesced along c
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            int c = c0 + r; long p = p0 + threadIdx.x;
            tile[r][threadIdx.x] = (c < C && p < P) ? xs[(long)c * P + p] : 0.f;
        }
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            long p = p0 + r; int c = c0 + threadIdx.x;
            if (p < P && c < C) ys[p * C + c] = tile[threadIdx.x][r];
        }
    } else {        // read [p][c] coalesced along c, write [c][p] coalesced along p
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            long p = p0 + r; int c = c0 + threadIdx.x;
            tile[r][threadIdx.x] = (p < P && c < C) ? xs[p * C + c] : 0.f;
        }
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            int c = c0 + r; long p = p0 + threadIdx.x;
            if (c < C && p < P) ys[(long)c * P + p] = tile[threadIdx.x][r];
        }
    }
}

// strided channel-block copy: dst[p][doff + c] (=|+=) src[p][soff + c], c < nC   (concat / slice)
// (division-free: one warp walks the channels of one pixel; for a handful of channels one lane owns one pixel)
__global__ void copy_channels_kernel(const float* __restrict__ src, int srcC, int soff, float* __restrict__ dst,
                                     int dstC, int doff, int nC, long npix, int accumulate) {
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
    if (nC <= 8) {
        for (long p = tid; p < npix; p += nthr) {
            const float* s = src + p * srcC + soff;
            float* d = dst + p * dstC + doff;
            for (int c = 0; c < nC; ++c) { if (accumulate) d[c] += s[c]; else d[c] = s[c]; }
        }
        return;
    }
    const int lane = threadIdx.x & 31;
    for (long p = tid >> 5; p < npix; p += nthr >> 5) {
        const float* s = src + p * srcC + soff;
        float* d = dst + p * dstC + doff;
        for (int c = lane; c < nC; c += 32) { if (accumulate) d[c] += s[c]; else d[c] = s[c]; }
    }
}
__global__ void copy_channels_vec4_kernel(const float4* __restrict__ src, int srcC4, int soff4, float4* __restrict__ dst,
                                          int dstC4, int doff4, int nC4, long npix) {
    // G lanes walk the float4 groups of one pixel (G = 8, 16 or 32 by channel count): no per-item divisions
    const int G = nC4 >= 32 ? 32 : (nC4 >= 16 ? 16 : 8), sh = nC4 >= 32 ? 5 : (nC4 >= 16 ? 4 : 3);
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
    const int lane = (int)(tid & (G - 1));
    for (long p = tid >> sh; p < npix; p += nthr >> sh) {
        const float4* s = src + p * srcC4 + soff4;
        float4* d = dst + p * dstC4 + doff4;
        for (int c = lane; c < nC4; c += G) d[c] = s[c];
    }
}

// ------------------------------------------------------------------------------------------
// padding (zero / reflect / replicate), NHWC
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int pad_src(int q, int p, int n, int mode) {  // padded index -> source index or -1
    int i = q - p;
    if (i >= 0 && i < n) return i;
    if (mode == DSR_PAD_ZERO) return -1;
    if (mode == DSR_PAD_REFLECT) return i < 0 ? -i : 2 * (n - 1) - i;
    return i < 0 ? 0 : n - 1;  // replicate
}
__global__ void pad2d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C,
                                 int p, int mode) {
    int Hp = H + 2 * p, Wp = W + 2 * p;
    long total = (long)N * Hp * Wp * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C); long t = idx / C;
        int qw = (int)(t % Wp); t /= Wp;
        int qh = (int)(t % Hp); int n = (int)(t / Hp);
        int i = pad_src(qh, p, H, mode), j = pad_src(qw, p, W, mode);
        y[idx] = (i < 0 || j < 0) ? 0.f : x[(((long)n * H + i) * W + j) * C + c];
    }
}
// list of padded indices that read source index i
__device__ __forceinline__ int pad_readers(int i, int p, int n, int mode, int* qs) {
    int k = 0;
    qs[k++] = i + p;
    if (mode == DSR_PAD_REFLECT) {
        if (i >= 1 && i <= p) qs[k++] = p - i;
        if (i <= n - 2 && i >= n - 1 - p) qs[k++] = 2 * (n - 1) - i + p;
    } else if (mode == DSR_PAD_REPLICATE) {
        if (i == 0) for (int q = 0; q < p; ++q) qs[k++] = q;
        if (i == n - 1) for (int q = n + p; q < n + 2 * p; ++q) qs[k++] = q;
    }
    return k;
}
__global__ void pad2d_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int N, int H, int W, int C,
                                 int p, int mode, int Wp, const float* __restrict__ add) {
    int Hp = H + 2 * p;
    long total = (long)N * H * W * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C); long t = idx / C;
        int j = (int)(t % W); t /= W;
        int i = (int)(t % H); int n = (int)(t / H);
        int qh[16], qw[16];
        int nh = pad_readers(i, p, H, mode, qh), nw = pad_readers(j, p, W, mode, qw);
        float acc = 0.f;
        for (int a = 0; a < nh; ++a)
            for (int b = 0; b < nw; ++b) acc += gy[(((long)n * Hp + qh[a]) * Wp + qw[b]) * C + c];
        gx[idx] = add ? acc + add[idx] : acc;
    }
}

// float4 version: one thread = 4 channels of one source pixel; interior pixels have exactly one reader
__global__ void __launch_bounds__(256)
pad2d_bwd_vec4_kernel(const float4* __restrict__ gy, float4* __restrict__ gx, int N, int H, int W, int C4, int p, int mode, int Wp,
                      const float4* __restrict__ add) {
    const int Hp = H + 2 * p;                      // Wp = row pitch of gy in pixels (>= W + 2 p)
    const int total = N * H * W * C4;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c = idx % C4; int t = idx / C4;
        const int j = t % W; t /= W;
        const int i = t % H; const int n = t / H;
        float4 acc;
        if (i > p && i < H - 1 - p && j > p && j < W - 1 - p) {
            acc = gy[((long)(n * Hp + i + p) * Wp + j + p) * C4 + c];
        } else {
            int qh[16], qw[16];
            const int nh = pad_readers(i, p, H, mode, qh), nw = pad_readers(j, p, W, mode, qw);
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int a = 0; a < nh; ++a)
                for (int b = 0; b < nw; ++b) {
                    const float4 v = gy[((long)(n * Hp + qh[a]) * Wp + qw[b]) * C4 + c];
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
        }
        if (add) {       // a second gradient of the same tensor (the skip connection of a residual block) joins in the same pass
            const float4 r = add[idx];
            acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
        }
        gx[idx] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// activations
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float x, int kind, float slope) {
    if (kind == DSR_ACT_RELU) return x > 0.f ? x : 0.f;
    if (kind == DSR_ACT_LRELU) return x > 0.f ? x : slope * x;
    if (kind == DSR_ACT_TANH) return tanhf(x);
    return x;
}
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long n, int kind, float slope) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        y[i] = act_apply(x[i], kind, slope);
}
// ref = x for relu / lrelu, = y for tanh
__global__ void act_bwd_kernel(const float* __restrict__ ref, const float* __restrict__ gy, float* __restrict__ gx,
                               long n, int kind, float slope) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float r = ref[i], g = gy[i];
        float d = 1.f;
        if (kind == DSR_ACT_RELU) d = r > 0.f ? 1.f : 0.f;
        else if (kind == DSR_ACT_LRELU) d = r > 0.f ? 1.f : slope;
        else if (kind == DSR_ACT_TANH) d = 1.f - r * r;
        gx[i] = g * d;
    }
}

// ------------------------------------------------------------------------------------------
// per-(n, c) sums over the pixels of an NHWC tensor:
//   mode 0: sums[n][c] = (sum x, sum x^2)                         (norm statistics, bias gradient)
//   mode 1: sums[n][c] = (sum dy', sum dy' * xhat), dy' = dy * 1[xhat > 0 if relu]   (IN backward)
// blockDim = 256; lanes run along channels (coalesced), the rest of the block along pixels.
// ------------------------------------------------------------------------------------------
__global__ void channel_sums_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                    const float* __restrict__ prm, int N, long P, int C, long chunk, int mode,
                                    int act, double* __restrict__ sums) {
    __shared__ double sh[2][TPB];
    int n = blockIdx.y;
    int Cw = 1;
    while (Cw < C && Cw < TPB) Cw <<= 1;
    int rows = TPB / Cw;
    int tx = threadIdx.x % Cw, ty = threadIdx.x / Cw;
    long p_begin = (long)blockIdx.x * chunk, p_end = p_begin + chunk;
    if (p_end > P) p_end = P;
    const float* xs = x + (long)n * P * C;
    const float* ds = dy ? dy + (long)n * P * C : nullptr;
    long NC = (long)N * C;
    for (int c0 = 0; c0 < C; c0 += Cw) {
        int c = c0 + tx;
        double a = 0.0, b = 0.0;
        if (c < C) {
            float mean = 0.f, scale = 1.f;
            if (mode == 1) { mean = prm[(long)n * C + c]; scale = prm[NC + (long)n * C + c]; }
            for (long p = p_begin + ty; p < p_end; p += rows) {
                float v = xs[p * C + c];
                if (mode == 0) { a += (double)v; b += (double)v * (double)v; }
                else {
                    float xh = (v - mean) * scale;
                    float g = ds[p * C + c];
                    if (act == DSR_ACT_RELU && !(xh > 0.f)) g = 0.f;
                    a += (double)g; b += (double)(g * xh);
                }
            }
        }
        sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b;
        __syncthreads();
        if (ty == 0 && c < C) {
            for (int r = 1; r < rows; ++r) { a += sh[0][r * Cw + tx]; b += sh[1][r * Cw + tx]; }
            atomicAdd(&sums[((long)n * C + c) * 2], a);
            atomicAdd(&sums[((long)n * C + c) * 2 + 1], b);
        }
        __syncthreads();
    }
}

// mode 0 with 16-byte loads: lanes along channel quads, the rest of the block along pixels.  The scalar kernel above adds
// every element into fp64 accumulators (two half-rate instructions per 4 bytes: 58 % of the HBM peak); here a thread
// sums runs of 16 pixels in fp32 and folds each run into its fp64 totals, so the statistics keep their fp64 accumulation
// across the ~10^4..10^5 pixels of a plane while the inner loop is fp32.
__global__ void __launch_bounds__(256)
channel_sums_vec4_kernel(const float4* __restrict__ x, int P, int C4, int chunk, double* __restrict__ sums) {
    __shared__ double sh_d[256 * 8];
    const int n = blockIdx.y;
    const int Cw = C4 < 256 ? C4 : 256, rows = 256 / Cw;
    const int tx = threadIdx.x % Cw, ty = threadIdx.x / Cw;
    const long base = (long)n * P * C4;
    const int p_begin = blockIdx.x * chunk, p_end = min(p_begin + chunk, P);
    for (int c0 = 0; c0 < C4; c0 += Cw) {
        const int cq = c0 + tx;
        double da[4] = {0.0, 0.0, 0.0, 0.0}, db[4] = {0.0, 0.0, 0.0, 0.0};
        if (cq < C4 && ty < rows) {
            for (int p0 = p_begin + ty; p0 < p_end; p0 += 16 * rows) {
                float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
                if (p0 + 15 * rows < p_end) {
                    // whole run inside the chunk: 8 independent 16-byte loads in flight per thread (a reduction has nothing
                    // else to hide the HBM latency behind - 4 in flight ran at 66 % of the HBM peak, r2a)
                    const float4* q = x + base + (long)p0 * C4 + cq;
                    const long st = (long)rows * C4;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float4 v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] = __ldg(q + (h * 8 + u) * st);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            a[0] += v[u].x; a[1] += v[u].y; a[2] += v[u].z; a[3] += v[u].w;
                            b[0] += v[u].x * v[u].x; b[1] += v[u].y * v[u].y; b[2] += v[u].z * v[u].z; b[3] += v[u].w * v[u].w;
                        }
                    }
                } else {
#pragma unroll 4
                    for (int u = 0; u < 16; ++u) {
                        const int p = p0 + u * rows;
                        if (p < p_end) {
                            const float4 v = x[base + (long)p * C4 + cq];
                            a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
                            b[0] += v.x * v.x; b[1] += v.y * v.y; b[2] += v.z * v.z; b[3] += v.w * v.w;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) { da[k] += (double)a[k]; db[k] += (double)b[k]; }
            }
        }
        double* mine = sh_d + (ty * Cw + tx) * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) { mine[k] = da[k]; mine[4 + k] = db[k]; }
        __syncthreads();
        if (ty == 0 && cq < C4) {
            for (int r = 1; r < rows; ++r) {
                const double* o = sh_d + (r * Cw + tx) * 8;
#pragma unroll
                for (int k = 0; k < 4; ++k) { da[k] += o[k]; db[k] += o[4 + k]; }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                atomicAdd(&sums[((long)n * C4 * 4 + cq * 4 + k) * 2], da[k]);
                atomicAdd(&sums[((long)n * C4 * 4 + cq * 4 + k) * 2 + 1], db[k]);
            }
        }
        __syncthreads();
    }
}

// sums -> per-(n,c) (mean, scale, shift):  y = (x - mean) * scale + shift
//   groups == 0: instance norm (biased variance over P), scale = rstd, shift = 0
//   groups  > 0: group norm over (P x C/groups), scale = rstd_g * gamma_c, shift = beta_c
__global__ void norm_finalize_kernel(const double* __restrict__ sums, int N, int C, long P, int groups,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                     float* __restrict__ prm) {
    long NC = (long)N * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < NC; idx += (long)gridDim.x * blockDim.x) {
        int n = (int)(idx / C), c = (int)(idx % C);
        float mean, scale, shift;
        norm_fin_one(sums, P, groups, gamma, beta, eps, n, c, C, mean, scale, shift);
        prm[idx] = mean;
        prm[NC + idx] = scale;
        prm[2 * NC + idx] = shift;
    }
}

// y = act((x - mean) * scale + shift) (+ residual)
__global__ void norm_apply_fwd_kernel(const float* __restrict__ x, const float* __restrict__ prm,
                                      const float* __restrict__ res, float* __restrict__ y, int N, long P, int C,
                                      int act) {
    long NC = (long)N * C, total = (long)N * P * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C);
        int n = (int)(idx / (P * C));
        long k = (long)n * C + c;
        float v = (x[idx] - prm[k]) * prm[NC + k] + prm[2 * NC + k];
        if (act == DSR_ACT_RELU) v = v > 0.f ? v : 0.f;
        if (res) v += res[idx];
        y[idx] = v;
    }
}
// instance-norm backward: dx = rstd * (dy' - mean(dy') - xhat * mean(dy' * xhat))
__global__ void in_apply_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                    const float* __restrict__ prm, const double* __restrict__ sums2,
                                    float* __restrict__ dx, int N, long P, int C, int act) {
    long NC = (long)N * C, total = (long)N * P * C;
    float invP = 1.f / (float)P;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C);
        int n = (int)(idx / (P * C));
        long k = (long)n * C + c;
        float rstd = prm[NC + k];
        float xh = (x[idx] - prm[k]) * rstd;
        float g = dy[idx];
        if (act == DSR_ACT_RELU && !(xh > 0.f)) g = 0.f;
        float m1 = (float)sums2[k * 2] * invP, m2 = (float)sums2[k * 2 + 1] * invP;
        dx[idx] = rstd * (g - m1 - xh * m2);
    }
}

// float4 versions (C % 4 == 0, 16-byte aligned): blockIdx.y = image, one thread-item = 4 channels of one pixel.
// These passes are pure HBM streams (12 .. 20 B per element); the scalar versions above spent their time in 64-bit
// div / mod per element.
__global__ void __launch_bounds__(256)
norm_apply_fwd_vec4_kernel(const float4* __restrict__ x, const float* __restrict__ prm, const float4* __restrict__ res,
                           float4* __restrict__ y, int N, int P, int C4, int act) {
    const int n = blockIdx.y;
    const long NC = (long)N * C4 * 4;
    const float* pm = prm + (long)n * C4 * 4;
    const long base = (long)n * P * C4;
    const int total = P * C4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = (i % C4) * 4;
        const float4 m = ld4(pm + c), sc = ld4(pm + NC + c), sh = ld4(pm + 2 * NC + c);
        float4 v = x[base + i];
        v.x = (v.x - m.x) * sc.x + sh.x; v.y = (v.y - m.y) * sc.y + sh.y;
        v.z = (v.z - m.z) * sc.z + sh.z; v.w = (v.w - m.w) * sc.w + sh.w;
        if (act == DSR_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (res) { const float4 r = res[base + i]; v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        y[base + i] = v;
    }
}
// the same pass with dsr_norm_finalize folded in: each block derives its sample's constants from the raw sums into shared
// memory (norm_fin_one: the finalize kernel's own arithmetic), block x = 0 of every sample writes them out for the backward pass
__global__ void __launch_bounds__(256, 4)
norm_apply_fwd_fin_vec4_kernel(const float4* __restrict__ x, const NormFin fin, const float4* __restrict__ res,
                               float4* __restrict__ y, int N, int P, int C4, int act) {
    extern __shared__ __align__(16) float s_fin[];          // [3][C]
    const int n = blockIdx.y, C = C4 * 4;
    const long NC = (long)N * C;
    for (int c = threadIdx.x; c < C; c += 256) {
        float m, sc, sh;
        norm_fin_one(fin.sums, fin.P, fin.groups, fin.gamma, fin.beta, fin.eps, n, c, C, m, sc, sh);
        s_fin[c] = m; s_fin[C + c] = sc; s_fin[2 * C + c] = sh;
        if (blockIdx.x == 0 && fin.prm_out) {
            fin.prm_out[(long)n * C + c] = m; fin.prm_out[NC + (long)n * C + c] = sc; fin.prm_out[2 * NC + (long)n * C + c] = sh;
        }
    }
    __syncthreads();
    const long base = (long)n * P * C4;
    const int total = P * C4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = (i % C4) * 4;
        const float4 m = ld4(s_fin + c), sc = ld4(s_fin + C + c), sh = ld4(s_fin + 2 * C + c);
        float4 v = x[base + i];
        v.x = (v.x - m.x) * sc.x + sh.x; v.y = (v.y - m.y) * sc.y + sh.y;
        v.z = (v.z - m.z) * sc.z + sh.z; v.w = (v.w - m.w) * sc.w + sh.w;
        if (act == DSR_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (res) { const float4 r = res[base + i]; v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        y[base + i] = v;
    }
}
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)).  The per-(n, c) constants (mean, rstd, and the two fp64 sums turned
// into fp32 means) are staged once per block in shared memory; the old form re-read 8 doubles and converted them per
// 16-byte item and ran at 39 % of HBM peak.
template <bool GN>
__global__ void __launch_bounds__(256)
in_apply_bwd_vec4_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, const float* __restrict__ prm,
                         const double* __restrict__ sums2, float4* __restrict__ dx, int N, int P, int C4, int act,
                         const float* __restrict__ gamma = nullptr, const float* __restrict__ beta = nullptr,
                         const float* __restrict__ coef = nullptr) {
    // GN: the GroupNorm form (affine gamma / beta, group means `coef` from gn_bwd_finalize): [6][C] adds gamma, beta
    extern __shared__ __align__(16) float s_c[];          // [4][C]: mean, rstd, m1, m2
    const int n = blockIdx.y, C = C4 * 4;
    const long NC = (long)N * C;
    const float invP = 1.f / (float)P;
    for (int c = threadIdx.x; c < C; c += 256) {
        s_c[c] = prm[(long)n * C + c];
        s_c[C + c] = prm[NC + (long)n * C + c];
        if (GN) {
            s_c[2 * C + c] = coef[((long)n * C + c) * 2];
            s_c[3 * C + c] = coef[((long)n * C + c) * 2 + 1];
            s_c[4 * C + c] = gamma[c];
            s_c[5 * C + c] = beta[c];
        } else {
            s_c[2 * C + c] = (float)sums2[((long)n * C + c) * 2] * invP;
            s_c[3 * C + c] = (float)sums2[((long)n * C + c) * 2 + 1] * invP;
        }
    }
    __syncthreads();
    const long base = (long)n * P * C4;
    const int total = P * C4;
    const bool pow2 = (C4 & (C4 - 1)) == 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = (pow2 ? (i & (C4 - 1)) : (i % C4)) * 4;
        const float4 m = *reinterpret_cast<const float4*>(s_c + c), rs = *reinterpret_cast<const float4*>(s_c + C + c);
        const float4 a1 = *reinterpret_cast<const float4*>(s_c + 2 * C + c), a2 = *reinterpret_cast<const float4*>(s_c + 3 * C + c);
        const float4 xv = x[base + i], g4 = dy[base + i];
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {g4.x, g4.y, g4.z, g4.w};
        const float ms[4] = {m.x, m.y, m.z, m.w}, rr[4] = {rs.x, rs.y, rs.z, rs.w};
        const float m1[4] = {a1.x, a1.y, a1.z, a1.w}, m2[4] = {a2.x, a2.y, a2.z, a2.w};
        float o[4];
        if (GN) {
            const float4 g4a = *reinterpret_cast<const float4*>(s_c + 4 * C + c), b4a = *reinterpret_cast<const float4*>(s_c + 5 * C + c);
            const float ga[4] = {g4a.x, g4a.y, g4a.z, g4a.w}, be[4] = {b4a.x, b4a.y, b4a.z, b4a.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float xh = (xs[k] - ms[k]) * rr[k];
                float g = gs[k];
                if (act == DSR_ACT_RELU && !(xh * ga[k] + be[k] > 0.f)) g = 0.f;
                o[k] = rr[k] * (ga[k] * g - m1[k] - xh * m2[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = in_bwd_one(xs[k], gs[k], ms[k], rr[k], m1[k], m2[k], act == DSR_ACT_RELU);
        }
        dx[base + i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}
__global__ void __launch_bounds__(256)
act_bwd_vec4_kernel(const float4* __restrict__ ref, const float4* __restrict__ gy, float4* __restrict__ gx, long n4, int kind,
                    float slope) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 r = ref[i], g = gy[i];
        const float rs[4] = {r.x, r.y, r.z, r.w}, gs[4] = {g.x, g.y, g.z, g.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float d = 1.f;
            if (kind == DSR_ACT_RELU) d = rs[k] > 0.f ? 1.f : 0.f;
            else if (kind == DSR_ACT_LRELU) d = rs[k] > 0.f ? 1.f : slope;
            else if (kind == DSR_ACT_TANH) d = 1.f - rs[k] * rs[k];
            o[k] = gs[k] * d;
        }
        gx[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}
// IN-backward sums, float4: lanes run along channel quads, warps along pixels; per-(n, c) (sum dy', sum dy' * xhat)
template <bool GN>      // GN: ReLU gate on xhat * gamma + beta (GroupNorm with affine parameters)
__global__ void __launch_bounds__(256)
in_bwd_sums_vec4_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, const float* __restrict__ prm, int N, int P,
                        int C4, int chunk, int act, double* __restrict__ sums, const float* __restrict__ gamma = nullptr,
                        const float* __restrict__ beta = nullptr) {
    extern __shared__ float sh_s[];              // [rows][C4*4][2]
    const int n = blockIdx.y;
    const int Cw = C4 < 256 ? C4 : 256;          // threads along channel quads (C4 <= 256 here)
    const int rows = 256 / Cw;
    const int tx = threadIdx.x % Cw, ty = threadIdx.x / Cw;
    const long NC = (long)N * C4 * 4;
    const float* pm = prm + (long)n * C4 * 4;
    const long base = (long)n * P * C4;
    int p_begin = blockIdx.x * chunk, p_end = p_begin + chunk;
    if (p_end > P) p_end = P;
    for (int c0 = 0; c0 < C4; c0 += Cw) {
        const int cq = c0 + tx;
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (cq < C4 && ty < rows) {
            const float4 m = ld4(pm + cq * 4), rs = ld4(pm + NC + cq * 4);
            const float ms[4] = {m.x, m.y, m.z, m.w}, rr[4] = {rs.x, rs.y, rs.z, rs.w};
            float ga[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
            if (GN) {
                const float4 g4 = ld4(gamma + cq * 4), b4 = ld4(beta + cq * 4);
                ga[0] = g4.x; ga[1] = g4.y; ga[2] = g4.z; ga[3] = g4.w; be[0] = b4.x; be[1] = b4.y; be[2] = b4.z; be[3] = b4.w;
            }
            for (int p = p_begin + ty; p < p_end; p += rows) {
                const float4 xv = x[base + (long)p * C4 + cq], gv = dy[base + (long)p * C4 + cq];
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float xh = (xs[k] - ms[k]) * rr[k];
                    float g = gs[k];
                    if (act == DSR_ACT_RELU && !((GN ? xh * ga[k] + be[k] : xh) > 0.f)) g = 0.f;
                    a[k] += g; b[k] += g * xh;
                }
            }
        }
        // reduce over the `rows` pixel lanes of the block, then one fp64 atomic per (n, c, quantity)
        float* mine = sh_s + ((ty * Cw + tx) * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) { mine[k] = a[k]; mine[4 + k] = b[k]; }
        __syncthreads();
        if (ty == 0 && cq < C4) {
            double da[4], db[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { da[k] = 0.0; db[k] = 0.0; }
            for (int r = 0; r < rows; ++r) {
                const float* o = sh_s + ((r * Cw + tx) * 8);
#pragma unroll
                for (int k = 0; k < 4; ++k) { da[k] += (double)o[k]; db[k] += (double)o[4 + k]; }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                atomicAdd(&sums[((long)n * C4 * 4 + cq * 4 + k) * 2], da[k]);
                atomicAdd(&sums[((long)n * C4 * 4 + cq * 4 + k) * 2 + 1], db[k]);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// weights: 4-D parameter [D0][D1][R][S] <-> packed GEMM operand [(r*S+s)*Ck + ck][Co]
//   kdim == 1: ck indexes D1, co indexes D0 (Conv2d forward, ConvTranspose2d dgrad)
//   kdim == 0: ck indexes D0, co indexes D1 (Conv2d dgrad, ConvTranspose2d forward)
// ------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int D0, int D1, int R, int S, int kdim,
                                   float* __restrict__ out) {
    long total = (long)D0 * D1 * R * S;
    int Ck = kdim ? D1 : D0, Co = kdim ? D0 : D1;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int co = (int)(idx % Co); long t = idx / Co;
        int ck = (int)(t % Ck); int tap = (int)(t / Ck);
        int d0 = kdim ? co : ck, d1 = kdim ? ck : co;
        out[idx] = w[((long)d0 * D1 + d1) * R * S + tap];
    }
}
__global__ void unpack_weight_kernel(const float* __restrict__ packed, int D0, int D1, int R, int S, int kdim,
                                     float* __restrict__ w, int accumulate) {
    long total = (long)D0 * D1 * R * S;
    int Ck = kdim ? D1 : D0, Co = kdim ? D0 : D1;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int tap = (int)(idx % (R * S)); long t = idx / (R * S);
        int d1 = (int)(t % D1); int d0 = (int)(t / D1);
        int ck = kdim ? d1 : d0, co = kdim ? d0 : d1;
        float v = packed[((long)tap * Ck + ck) * Co + co];
        if (accumulate) w[idx] += v; else w[idx] = v;
    }
}

// double -> float with scale (bias gradient, loss means)
__global__ void cvt_f64_f32_kernel(const double* __restrict__ in, long stride_in, float* __restrict__ out, long n,
                                   float scale, int accumulate) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float v = (float)(in[i * stride_in] * (double)scale);
        if (accumulate) out[i] += v; else out[i] = v;
    }
}

// out (+)= sum over `reps` replica rows of a double accumulator (the replicated bias-gradient sums of dsr_tc_prep)
__global__ void sum_reps_f64_f32_kernel(const double* __restrict__ in, int reps, long n, float* __restrict__ out, int accumulate) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double a = 0;
        for (int r = 0; r < reps; ++r) a += in[(long)r * n + i];
        if (accumulate) out[i] += (float)a; else out[i] = (float)a;
    }
}

// ------------------------------------------------------------------------------------------
// Adam over one flat fp32 arena (torch.optim.Adam semantics, no weight decay, no amsgrad)
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float step, float b2, float omb1, float omb2,
                            float eps, float bc2_sqrt, float grad_scale) {
    // omb1 = 1 - beta1, omb2 = 1 - beta2 and step = lr / (1 - beta1^t) are formed in double on the host,
    // like torch.optim.Adam does (1 - 0.999f in fp32 is off by 4.7e-5 relative)
    long n4 = n >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 P = ld4(p + 4 * i), G = ld4(g + 4 * i), M = ld4(m + 4 * i), V = ld4(v + 4 * i);
        float* pp = &P.x; float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gr = gg[k] * grad_scale;
            mm[k] = mm[k] + (gr - mm[k]) * omb1;
            vv[k] = vv[k] * b2 + gr * gr * omb2;
            float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
            pp[k] -= step * (mm[k] / denom);
        }
        st4(p + 4 * i, P); st4(m + 4 * i, M); st4(v + 4 * i, V);
    }
    if (blockIdx.x == 0) {
        for (long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            float gr = g[i] * grad_scale;
            float mk = m[i] + (gr - m[i]) * omb1;
            float vk = v[i] * b2 + gr * gr * omb2;
            m[i] = mk; v[i] = vk;
            p[i] -= step * (mk / (sqrtf(vk) / bc2_sqrt + eps));
        }
    }
}

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
extern "C" int dsr_nchw_to_nhwc(const float* x, float* y, int N, int C, long P, void* stream) {
    DSR_REQUIRE(x && y && N > 0 && C > 0 && P > 0, "bad arguments");
    dim3 grid(dsr_cdiv(P, 32), dsr_cdiv(C, 32), N), block(32, 8);
    transpose_cp_kernel<<<grid, block, 0, ST(stream)>>>(x, y, C, P, 1);
    return dsr_check_launch("nchw_to_nhwc");
}
extern "C" int dsr_nhwc_to_nchw(const float* x, float* y, int N, int C, long P, void* stream) {
    DSR_REQUIRE(x && y && N > 0 && C > 0 && P > 0, "bad arguments");
    dim3 grid(dsr_cdiv(P, 32), dsr_cdiv(C, 32), N), block(32, 8);
    transpose_cp_kernel<<<grid, block, 0, ST(stream)>>>(x, y, C, P, 0);
    return dsr_check_launch("nhwc_to_nchw");
}
extern "C" int dsr_copy_channels(const float* src, int srcC, int soff, float* dst, int dstC, int doff, int nC,
                                 long npix, int accumulate, void* stream) {
    DSR_REQUIRE(src && dst && nC > 0 && soff + nC <= srcC && doff + nC <= dstC, "bad channel ranges");
    bool vec = !accumulate && !(srcC & 3) && !(dstC & 3) && !(soff & 3) && !(doff & 3) && !(nC & 3) &&
               !((uintptr_t)src & 15) && !((uintptr_t)dst & 15);
    if (vec)
        copy_channels_vec4_kernel<<<dsr_grid(npix * (nC / 4 >= 32 ? 32 : (nC / 4 >= 16 ? 16 : 8)), TPB), TPB, 0, ST(stream)>>>(
            (const float4*)src, srcC / 4, soff / 4, (float4*)dst, dstC / 4, doff / 4, nC / 4, npix);
    else
        copy_channels_kernel<<<dsr_grid(nC <= 8 ? npix : npix * 32, TPB), TPB, 0, ST(stream)>>>(src, srcC, soff, dst, dstC, doff, nC,
                                                                                                 npix, accumulate);
    return dsr_check_launch("copy_channels");
}
extern "C" int dsr_pad2d_fwd(const float* x, float* y, int N, int H, int W, int C, int pad, int mode, void* stream) {
    DSR_REQUIRE(x && y && pad >= 0 && pad <= 7, "bad arguments");
    DSR_REQUIRE(mode != DSR_PAD_REFLECT || (pad < H && pad < W), "reflect padding needs pad < size");
    pad2d_fwd_kernel<<<dsr_grid((long)N * (H + 2 * pad) * (W + 2 * pad) * C, TPB), TPB, 0, ST(stream)>>>(x, y, N, H, W, C,
                                                                                                     pad, mode);
    return dsr_check_launch("pad2d_fwd");
}
// wpitch = row pitch of gy in pixels (>= W + 2 pad): the grouped data gradient (ops._tc_dgrad_group) computes rows rounded up
// to whole pixel groups
extern "C" int dsr_pad2d_bwd_pitch_add(const float* gy, const float* add, float* gx, int N, int H, int W, int C, int pad, int mode,
                                       int wpitch, void* stream) {
    DSR_REQUIRE(gy && gx && pad >= 0 && pad <= 7 && wpitch >= W + 2 * pad, "bad arguments");
    if (!(C & 3) && !((uintptr_t)gy & 15) && !((uintptr_t)gx & 15) && !((uintptr_t)add & 15) && (long)N * H * W * (C / 4) < (1L << 31) &&
        (long)N * (H + 2 * pad) * wpitch * (C / 4) < (1L << 31))
        pad2d_bwd_vec4_kernel<<<dsr_grid((long)N * H * W * (C / 4), 256), 256, 0, ST(stream)>>>((const float4*)gy, (float4*)gx, N, H, W,
                                                                                               C / 4, pad, mode, wpitch, (const float4*)add);
    else
        pad2d_bwd_kernel<<<dsr_grid((long)N * H * W * C, TPB), TPB, 0, ST(stream)>>>(gy, gx, N, H, W, C, pad, mode, wpitch, add);
    return dsr_check_launch("pad2d_bwd");
}
extern "C" int dsr_pad2d_bwd_pitch(const float* gy, float* gx, int N, int H, int W, int C, int pad, int mode, int wpitch,
                                   void* stream) {
    return dsr_pad2d_bwd_pitch_add(gy, nullptr, gx, N, H, W, C, pad, mode, wpitch, stream);
}
extern "C" int dsr_pad2d_bwd(const float* gy, float* gx, int N, int H, int W, int C, int pad, int mode, void* stream) {
    return dsr_pad2d_bwd_pitch(gy, gx, N, H, W, C, pad, mode, W + 2 * pad, stream);
}
extern "C" int dsr_act_fwd(const float* x, float* y, long n, int kind, float slope, void* stream) {
    DSR_REQUIRE(x && y, "null pointer");
    act_fwd_kernel<<<dsr_grid(n, TPB), TPB, 0, ST(stream)>>>(x, y, n, kind, slope);
    return dsr_check_launch("act_fwd");
}
extern "C" int dsr_act_bwd(const float* ref, const float* gy, float* gx, long n, int kind, float slope, void* stream) {
    DSR_REQUIRE(ref && gy && gx, "null pointer");
    if (!(n & 3) && !((uintptr_t)ref & 15) && !((uintptr_t)gy & 15) && !((uintptr_t)gx & 15))
        act_bwd_vec4_kernel<<<dsr_grid(n / 4, 256), 256, 0, ST(stream)>>>((const float4*)ref, (const float4*)gy, (float4*)gx, n / 4, kind,
                                                                         slope);
    else
        act_bwd_kernel<<<dsr_grid(n, TPB), TPB, 0, ST(stream)>>>(ref, gy, gx, n, kind, slope);
    return dsr_check_launch("act_bwd");
}
static void sums_launch_cfg(int N, long P, long* chunk, dim3* grid) {
    long want = (long)dsr_num_sms() * 4 / (N > 0 ? N : 1);
    if (want < 1) want = 1;
    long c = (P + want - 1) / want;
    if (c < 64) c = 64;
    *chunk = c;
    *grid = dim3((unsigned)((P + c - 1) / c), (unsigned)N);
}
extern "C" int dsr_channel_sums(const float* x, int N, long P, int C, double* sums, void* stream) {
    DSR_REQUIRE(x && sums && N > 0 && P > 0 && C > 0, "bad arguments");
    long chunk; dim3 grid;
    sums_launch_cfg(N, P, &chunk, &grid);
    if (!(C & 3) && C / 4 <= 256 && (256 % (C / 4)) == 0 && !((uintptr_t)x & 15) && P * (C / 4) < (1L << 31))
        channel_sums_vec4_kernel<<<grid, 256, 0, ST(stream)>>>((const float4*)x, (int)P, C / 4, (int)chunk, sums);
    else
        channel_sums_kernel<<<grid, TPB, 0, ST(stream)>>>(x, nullptr, nullptr, N, P, C, chunk, 0, 0, sums);
    return dsr_check_launch("channel_sums");
}
extern "C" int dsr_norm_finalize(const double* sums, int N, int C, long P, int groups, const float* gamma,
                                 const float* beta, float eps, float* prm, void* stream) {
    DSR_REQUIRE(sums && prm && (groups == 0 || C % groups == 0), "bad arguments");
    norm_finalize_kernel<<<dsr_grid((long)N * C, TPB), TPB, 0, ST(stream)>>>(sums, N, C, P, groups, gamma, beta, eps, prm);
    return dsr_check_launch("norm_finalize");
}
extern "C" int dsr_norm_apply_fwd(const float* x, const float* prm, const float* res, float* y, int N, long P, int C,
                                  int act, void* stream) {
    DSR_REQUIRE(x && prm && y, "null pointer");
    const bool vec = !(C & 3) && !((uintptr_t)x & 15) && !((uintptr_t)y & 15) && !((uintptr_t)res & 15) && !((uintptr_t)prm & 15) &&
                     !((N * (long)C) & 3) && P * (C / 4) < (1L << 31) && C <= 2048;       // 4 C floats of shared memory
    if (vec) {
        const long items = P * (C / 4);
        long gx = (items + 255) / 256, cap = (long)dsr_num_sms() * 8 / N + 1;
        if (gx > cap) gx = cap;
        norm_apply_fwd_vec4_kernel<<<dim3((unsigned)gx, (unsigned)N), 256, 0, ST(stream)>>>((const float4*)x, prm, (const float4*)res,
                                                                                           (float4*)y, N, (int)P, C / 4, act);
    } else {
        norm_apply_fwd_kernel<<<dsr_grid((long)N * P * C, TPB), TPB, 0, ST(stream)>>>(x, prm, res, y, N, P, C, act);
    }
    return dsr_check_launch("norm_apply_fwd");
}
// dsr_norm_finalize + dsr_norm_apply_fwd in one launch (prm_out = what dsr_norm_finalize would have written, bit-identical)
extern "C" int dsr_norm_apply_fwd_fin(const float* x, const double* sums, int groups, const float* gamma, const float* beta,
                                      float eps, float* prm_out, const float* res, float* y, int N, long P, int C, int act,
                                      void* stream) {
    DSR_REQUIRE(x && sums && prm_out && y && N > 0 && C > 0 && P > 0, "null pointer / bad shape");
    DSR_REQUIRE(groups >= 0 && (groups == 0 || C % groups == 0), "C must be a multiple of groups");
    const bool vec = !(C & 3) && !((uintptr_t)x & 15) && !((uintptr_t)y & 15) && !((uintptr_t)res & 15) &&
                     P * (C / 4) < (1L << 31) && C <= 2048;
    if (!vec) {
        int rc = dsr_norm_finalize(sums, N, C, P, groups, gamma, beta, eps, prm_out, stream);
        return rc ? rc : dsr_norm_apply_fwd(x, prm_out, res, y, N, P, C, act, stream);
    }
    NormFin fin;
    fin.sums = sums; fin.gamma = gamma; fin.beta = beta; fin.prm_out = prm_out; fin.P = P; fin.groups = groups; fin.eps = eps;
    const long items = P * (C / 4);
    long gx = (items + 255) / 256, cap = (long)dsr_num_sms() * 8 / N + 1;
    if (gx > cap) gx = cap;
    norm_apply_fwd_fin_vec4_kernel<<<dim3((unsigned)gx, (unsigned)N), 256, 3 * (size_t)C * sizeof(float), ST(stream)>>>(
        (const float4*)x, fin, (const float4*)res, (float4*)y, N, (int)P, C / 4, act);
    return dsr_check_launch("norm_apply_fwd_fin");
}
// ------------------------------------------------------------------------------------------
// GroupNorm(groups, C, affine) [+ReLU] backward (translation_network.py:46; the generators of the translation block).
//   xhat = (x - mean_g) * rstd_g,  y = act(xhat * gamma + beta),  g = dy * act'(.)
//   dbeta_c = sum g, dgamma_c = sum g * xhat,
//   dx = rstd_g * (gamma_c * g - mean_g(gamma * g) - xhat * mean_g(gamma * g * xhat))      (means over the group's P * C/G elements)
// prm = (mean, rstd, 0) per (n, c) as written by norm_finalize with gamma = beta = NULL.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_bwd_sums_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ prm,
                   const float* __restrict__ gamma, const float* __restrict__ beta, int N, long P, int C, int act,
                   long chunk, double* __restrict__ sums2) {
    extern __shared__ float gsm[];                 // [2][C]
    const int n = blockIdx.y;
    const long NC = (long)N * C, p0 = (long)blockIdx.x * chunk, p1 = p0 + chunk < P ? p0 + chunk : P;
    for (int c = threadIdx.x; c < 2 * C; c += 256) gsm[c] = 0.f;
    __syncthreads();
    const long base = (long)n * P * C;
    for (long i = p0 * C + threadIdx.x; i < p1 * C; i += 256) {
        const int c = (int)(i % C);
        const long k = (long)n * C + c;
        const float xh = (x[base + i] - prm[k]) * prm[NC + k];
        float g = dy[base + i];
        if (act == DSR_ACT_RELU && !(xh * gamma[c] + beta[c] > 0.f)) g = 0.f;
        atomicAdd(&gsm[c], g);
        atomicAdd(&gsm[C + c], g * xh);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        atomicAdd(&sums2[((long)n * C + c) * 2], (double)gsm[c]);
        atomicAdd(&sums2[((long)n * C + c) * 2 + 1], (double)gsm[C + c]);
    }
}
// per (n, c): the two group means the apply pass needs; per c: dgamma / dbeta (=|+=)
__global__ void gn_bwd_finalize_kernel(const double* __restrict__ sums2, const float* __restrict__ gamma, int N, int C, int groups,
                                       long P, float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       int accumulate) {
    const int cg = C / groups;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long)N * C; idx += (long)gridDim.x * blockDim.x) {
        const int n = (int)(idx / C), c = (int)(idx % C), g0 = (c / cg) * cg;
        double a = 0, b = 0;
        for (int k = 0; k < cg; ++k) {
            a += (double)gamma[g0 + k] * sums2[((long)n * C + g0 + k) * 2];
            b += (double)gamma[g0 + k] * sums2[((long)n * C + g0 + k) * 2 + 1];
        }
        const double M = (double)P * cg;
        coef[idx * 2] = (float)(a / M);
        coef[idx * 2 + 1] = (float)(b / M);
        if (n == 0) {
            double s1 = 0, s2 = 0;
            for (int m = 0; m < N; ++m) { s1 += sums2[((long)m * C + c) * 2]; s2 += sums2[((long)m * C + c) * 2 + 1]; }
            if (accumulate) { dbeta[c] += (float)s1; dgamma[c] += (float)s2; } else { dbeta[c] = (float)s1; dgamma[c] = (float)s2; }
        }
    }
}
__global__ void gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ prm,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ coef,
                                    float* __restrict__ dx, int N, long P, int C, int act) {
    const long NC = (long)N * C, total = (long)N * P * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const int n = (int)(idx / (P * C));
        const long k = (long)n * C + c;
        const float rstd = prm[NC + k], xh = (x[idx] - prm[k]) * rstd;
        float g = dy[idx];
        if (act == DSR_ACT_RELU && !(xh * gamma[c] + beta[c] > 0.f)) g = 0.f;
        dx[idx] = rstd * (gamma[c] * g - coef[k * 2] - xh * coef[k * 2 + 1]);
    }
}
extern "C" int dsr_gn_bwd_sums(const float* x, const float* dy, const float* prm, const float* gamma, const float* beta, int N,
                               long P, int C, int act, double* sums2, void* stream) {
    DSR_REQUIRE(x && dy && prm && gamma && beta && sums2 && N > 0 && N <= 65535 && P > 0 && C > 0 && C <= 4096, "bad arguments");
    const bool vec = !(C & 3) && C / 4 <= 256 && (256 % (C / 4)) == 0 && !((uintptr_t)x & 15) && !((uintptr_t)dy & 15) &&
                     !((uintptr_t)prm & 15) && !((uintptr_t)gamma & 15) && !((uintptr_t)beta & 15) && !((N * (long)C) & 3) &&
                     P * (C / 4) < (1L << 31);
    if (vec) {              // float4 loads, register accumulation, one fp64 atomic per (block, channel, quantity)
        long vchunk; dim3 grid;
        sums_launch_cfg(N, P, &vchunk, &grid);
        in_bwd_sums_vec4_kernel<true><<<grid, 256, 256 * 8 * sizeof(float), ST(stream)>>>((const float4*)x, (const float4*)dy, prm, N, (int)P,
                                                                                         C / 4, (int)vchunk, act, sums2, gamma, beta);
        return dsr_check_launch("gn_bwd_sums");
    }
    long blocks = (long)dsr_num_sms() * 4 / N + 1;
    if (blocks > P) blocks = P;
    const long chunk = (P + blocks - 1) / blocks;
    gn_bwd_sums_kernel<<<dim3((unsigned)((P + chunk - 1) / chunk), (unsigned)N), 256, 2 * (size_t)C * sizeof(float), ST(stream)>>>(
        x, dy, prm, gamma, beta, N, P, C, act, chunk, sums2);
    return dsr_check_launch("gn_bwd_sums");
}
extern "C" int dsr_gn_bwd_finalize(const double* sums2, const float* gamma, int N, int C, int groups, long P, float* coef,
                                   float* dgamma, float* dbeta, int accumulate, void* stream) {
    DSR_REQUIRE(sums2 && gamma && coef && dgamma && dbeta && groups > 0 && C % groups == 0, "bad arguments");
    gn_bwd_finalize_kernel<<<dsr_grid((long)N * C, TPB), TPB, 0, ST(stream)>>>(sums2, gamma, N, C, groups, P, coef, dgamma, dbeta, accumulate);
    return dsr_check_launch("gn_bwd_finalize");
}
extern "C" int dsr_gn_bwd_apply(const float* x, const float* dy, const float* prm, const float* gamma, const float* beta,
                                const float* coef, float* dx, int N, long P, int C, int act, void* stream) {
    DSR_REQUIRE(x && dy && prm && gamma && beta && coef && dx, "null pointer");
    const bool vec = !(C & 3) && !((uintptr_t)x & 15) && !((uintptr_t)dy & 15) && !((uintptr_t)dx & 15) && !((uintptr_t)prm & 15) &&
                     !((N * (long)C) & 3) && P * (C / 4) < (1L << 31) && C <= 1024;       // 6 C floats of shared memory
    if (vec) {
        const long items = P * (C / 4);
        long gx = (items + 255) / 256, cap = (long)dsr_num_sms() * 8 / N + 1;
        if (gx > cap) gx = cap;
        in_apply_bwd_vec4_kernel<true><<<dim3((unsigned)gx, (unsigned)N), 256, 6 * (size_t)C * sizeof(float), ST(stream)>>>(
            (const float4*)x, (const float4*)dy, prm, nullptr, (float4*)dx, N, (int)P, C / 4, act, gamma, beta, coef);
        return dsr_check_launch("gn_bwd_apply");
    }
    gn_bwd_apply_kernel<<<dsr_grid((long)N * P * C, TPB), TPB, 0, ST(stream)>>>(x, dy, prm, gamma, beta, coef, dx, N, P, C, act);
    return dsr_check_launch("gn_bwd_apply");
}
extern "C" int dsr_in_bwd_sums(const float* x, const float* dy, const float* prm, int N, long P, int C, int act,
                               double* sums2, void* stream) {
    DSR_REQUIRE(x && dy && prm && sums2, "null pointer");
    long chunk; dim3 grid;
    sums_launch_cfg(N, P, &chunk, &grid);
    const bool vec = !(C & 3) && C / 4 <= 256 && (256 % (C / 4 < 256 ? C / 4 : 256)) == 0 && !((uintptr_t)x & 15) && !((uintptr_t)dy & 15) &&
                     !((uintptr_t)prm & 15) && !((N * (long)C) & 3) && P * (C / 4) < (1L << 31);
    if (vec)
        in_bwd_sums_vec4_kernel<false><<<grid, 256, 256 * 8 * sizeof(float), ST(stream)>>>((const float4*)x, (const float4*)dy, prm, N, (int)P,
                                                                                          C / 4, (int)chunk, act, sums2);
    else
        channel_sums_kernel<<<grid, TPB, 0, ST(stream)>>>(x, dy, prm, N, P, C, chunk, 1, act, sums2);
    return dsr_check_launch("in_bwd_sums");
}
extern "C" int dsr_in_bwd_apply(const float* x, const float* dy, const float* prm, const double* sums2, float* dx,
                                int N, long P, int C, int act, void* stream) {
    DSR_REQUIRE(x && dy && prm && sums2 && dx, "null pointer");
    const bool vec = !(C & 3) && !((uintptr_t)x & 15) && !((uintptr_t)dy & 15) && !((uintptr_t)dx & 15) && !((uintptr_t)prm & 15) &&
                     !((N * (long)C) & 3) && P * (C / 4) < (1L << 31) && C <= 2048;       // 4 C floats of shared memory
    if (vec) {
        const long items = P * (C / 4);
        long gx = (items + 255) / 256, cap = (long)dsr_num_sms() * 8 / N + 1;
        if (gx > cap) gx = cap;
        in_apply_bwd_vec4_kernel<false><<<dim3((unsigned)gx, (unsigned)N), 256, 4 * (size_t)C * sizeof(float), ST(stream)>>>((const float4*)x, (const float4*)dy, prm, sums2,
                                                                                         (float4*)dx, N, (int)P, C / 4, act);
    } else {
        in_apply_bwd_kernel<<<dsr_grid((long)N * P * C, TPB), TPB, 0, ST(stream)>>>(x, dy, prm, sums2, dx, N, P, C, act);
    }
    return dsr_check_launch("in_bwd_apply");
}
extern "C" int dsr_pack_weight(const float* w, int D0, int D1, int R, int S, int kdim, float* out, void* stream) {
    DSR_REQUIRE(w && out, "null pointer");
    pack_weight_kernel<<<dsr_grid((long)D0 * D1 * R * S, TPB), TPB, 0, ST(stream)>>>(w, D0, D1, R, S, kdim, out);
    return dsr_check_launch("pack_weight");
}
extern "C" int dsr_unpack_weight(const float* packed, int D0, int D1, int R, int S, int kdim, float* w, int accumulate,
                                 void* stream) {
    DSR_REQUIRE(packed && w, "null pointer");
    unpack_weight_kernel<<<dsr_grid((long)D0 * D1 * R * S, TPB), TPB, 0, ST(stream)>>>(packed, D0, D1, R, S, kdim, w,
                                                                                         accumulate);
    return dsr_check_launch("unpack_weight");
}
extern "C" int dsr_cvt_f64_f32(const double* in, long stride_in, float* out, long n, float scale, int accumulate,
                               void* stream) {
    DSR_REQUIRE(in && out, "null pointer");
    cvt_f64_f32_kernel<<<dsr_grid(n, TPB), TPB, 0, ST(stream)>>>(in, stride_in, out, n, scale, accumulate);
    return dsr_check_launch("cvt_f64_f32");
}
extern "C" int dsr_sum_reps_f64_f32(const double* in, int reps, long n, float* out, int accumulate, void* stream) {
    DSR_REQUIRE(in && out && reps >= 1 && n >= 1, "bad arguments");
    sum_reps_f64_f32_kernel<<<dsr_grid(n, TPB), TPB, 0, ST(stream)>>>(in, reps, n, out, accumulate);
    return dsr_check_launch("sum_reps_f64_f32");
}
// Adam whose step count and hyper-parameters live on the device, so that the launch parameters never change and the
// whole training step can be replayed as a CUDA graph: `tick` advances the counter, every block of the update kernel
// derives the bias corrections from it (double precision, like torch.optim.Adam on the host).
__global__ void adam_tick_kernel(int* step) { if (threadIdx.x == 0 && blockIdx.x == 0) *step += 1; }
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long n, const double* __restrict__ hyper, const int* __restrict__ step,
                                float grad_scale) {
    __shared__ float sc[6];
    if (threadIdx.x == 0) {
        const double lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
        const int t = *step;
        sc[0] = (float)(lr / (1.0 - pow(b1, (double)t)));
        sc[1] = (float)b2; sc[2] = (float)(1.0 - b1); sc[3] = (float)(1.0 - b2); sc[4] = (float)eps;
        sc[5] = (float)sqrt(1.0 - pow(b2, (double)t));
    }
    __syncthreads();
    const float stepsz = sc[0], b2 = sc[1], omb1 = sc[2], omb2 = sc[3], eps = sc[4], bc2_sqrt = sc[5];
    long n4 = n >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 P = ld4(p + 4 * i), G = ld4(g + 4 * i), M = ld4(m + 4 * i), V = ld4(v + 4 * i);
        float* pp = &P.x; float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gr = gg[k] * grad_scale;
            mm[k] = mm[k] + (gr - mm[k]) * omb1;
            vv[k] = vv[k] * b2 + gr * gr * omb2;
            float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
            pp[k] -= stepsz * (mm[k] / denom);
        }
        st4(p + 4 * i, P); st4(m + 4 * i, M); st4(v + 4 * i, V);
    }
    if (blockIdx.x == 0) {
        for (long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            float gr = g[i] * grad_scale;
            float mk = m[i] + (gr - m[i]) * omb1;
            float vk = v[i] * b2 + gr * gr * omb2;
            m[i] = mk; v[i] = vk;
            p[i] -= stepsz * (mk / (sqrtf(vk) / bc2_sqrt + eps));
        }
    }
}
extern "C" int dsr_adam_step_dev(float* p, const float* g, float* m, float* v, long n, const double* hyper, int* step,
                                 float grad_scale, void* stream) {
    DSR_REQUIRE(p && g && m && v && hyper && step && n > 0, "bad arguments");
    DSR_REQUIRE(!((uintptr_t)p & 15) && !((uintptr_t)g & 15) && !((uintptr_t)m & 15) && !((uintptr_t)v & 15),
                "arena pointers must be 16-byte aligned");
    adam_tick_kernel<<<1, 32, 0, ST(stream)>>>(step);
    adam_dev_kernel<<<dsr_grid(n / 4 + 1, TPB), TPB, 0, ST(stream)>>>(p, g, m, v, n, hyper, step, grad_scale);
    return dsr_check_launch("adam_step_dev");
}
extern "C" int dsr_adam_step(float* p, const float* g, float* m, float* v, long n, double lr, double b1, double b2,
                             double eps, int step, float grad_scale, void* stream) {
    DSR_REQUIRE(p && g && m && v && n > 0 && step >= 1, "bad arguments");
    DSR_REQUIRE(!((uintptr_t)p & 15) && !((uintptr_t)g & 15) && !((uintptr_t)m & 15) && !((uintptr_t)v & 15),
                "arena pointers must be 16-byte aligned");
    // hyper-parameters arrive as doubles (python floats) so 1 - beta is formed exactly like torch.optim.Adam forms it
    double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
    adam_kernel<<<dsr_grid(n / 4 + 1, TPB), TPB, 0, ST(stream)>>>(
        p, g, m, v, n, (float)(lr / bc1), (float)b2, (float)(1.0 - b1), (float)(1.0 - b2), (float)eps,
        (float)sqrt(bc2), grad_scale);
    return dsr_check_launch("adam_step");
}

// ------------------------------------------------------------------------------------------
// loss assembly: loss_G = scale * sum_k sum_j w[k][j] * term_k[j]   (main_model.py:393-417) - one launch forward, one backward
// ------------------------------------------------------------------------------------------
#define LOSS_MAX_TERMS 32
struct LossTerms {
    const float* p[LOSS_MAX_TERMS];
    float w[LOSS_MAX_TERMS][4];
    int cnt[LOSS_MAX_TERMS];
    int n;
};
__global__ void loss_sum_fwd_kernel(const LossTerms t, float scale, float* __restrict__ out, int* __restrict__ nonfinite) {
    float a = 0.f;
    const int k = threadIdx.x;
    if (k < t.n)
        for (int j = 0; j < t.cnt[k]; ++j) a += t.w[k][j] * t.p[k][j];
    a = warp_sum(a);
    if (k == 0) {
        *out = a * scale;
        if (nonfinite && !isfinite(a * scale)) atomicAdd(nonfinite, 1);
    }
}
__global__ void loss_sum_bwd_kernel(const float* __restrict__ g, const LossTerms t, float scale, float* __restrict__ grads) {
    const int i = threadIdx.x, k = i >> 2, j = i & 3;
    if (k < t.n) grads[i] = *g * scale * t.w[k][j];
}
extern "C" int dsr_loss_sum_fwd(const float* const* terms, const int* counts, const float* weights, int n, float scale, float* out,
                                int* nonfinite, void* stream) {
    DSR_REQUIRE(terms && counts && weights && out && n > 0 && n <= LOSS_MAX_TERMS, "1..32 terms");
    LossTerms t;
    t.n = n;
    for (int k = 0; k < LOSS_MAX_TERMS; ++k) {
        t.p[k] = k < n ? terms[k] : nullptr;
        t.cnt[k] = k < n ? counts[k] : 0;
        for (int j = 0; j < 4; ++j) t.w[k][j] = k < n ? weights[4 * k + j] : 0.f;
        DSR_REQUIRE(k >= n || (t.p[k] && t.cnt[k] >= 1 && t.cnt[k] <= 4), "every term needs a pointer and 1..4 elements");
    }
    loss_sum_fwd_kernel<<<1, 32, 0, ST(stream)>>>(t, scale, out, nonfinite);
    return dsr_check_launch("loss_sum_fwd");
}
extern "C" int dsr_loss_sum_bwd(const float* g, const float* weights, int n, float scale, float* grads, void* stream) {
    DSR_REQUIRE(g && weights && grads && n > 0 && n <= LOSS_MAX_TERMS, "1..32 terms");
    LossTerms t;
    t.n = n;
    for (int k = 0; k < LOSS_MAX_TERMS; ++k) {
        t.p[k] = nullptr; t.cnt[k] = 0;
        for (int j = 0; j < 4; ++j) t.w[k][j] = k < n ? weights[4 * k + j] : 0.f;
    }
    loss_sum_bwd_kernel<<<1, 4 * LOSS_MAX_TERMS, 0, ST(stream)>>>(g, t, scale, grads);
    return dsr_check_launch("loss_sum_bwd");
}
