// Resampling kernels of the super-resolution path (models/main_sr_model.py:279-293, :361, :394-398, :452, :459):
//   F.interpolate(mode='bicubic')  (align_corners=False, A = -0.75, border-clamped taps)  forward + adjoint
//   F.interpolate(mode='nearest')  (src = min(floor(dst * in/out), in - 1))               forward
// Tensors are [N][H][W][C] (C = 1 for NCHW planes with N = B*C; C = 128 for the NHWC feature maps), so one kernel
// serves the depth / image planes and the channels-last activations.  Bandwidth bound: every output element is
// written once with coalesced (float4 when C % 4 == 0) accesses; the 16 taps of neighbouring outputs hit L1/L2.
#include "common.cuh"
#include "../../include/dsr_b200.h"

#define ST(s) ((cudaStream_t)(s))
static const int TPB = 256;

// cubic convolution coefficients for the 4 taps around floor(x) (Keys, A = -0.75), fraction t in [0, 1)
__device__ __forceinline__ void cubic_coeffs(float t, float w[4]) {
    const float A = -0.75f;
    float x = t + 1.f;
    w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
    x = t;
    w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    x = 1.f - t;
    w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    x = 2.f - t;
    w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}
// source coordinate of output index o: scale * (o + 0.5) - 0.5, NOT clamped at zero for the cubic filter
__device__ __forceinline__ void cubic_src(int o, float scale, int& i0, float w[4]) {
    const float r = scale * ((float)o + 0.5f) - 0.5f;
    const float f = floorf(r);
    i0 = (int)f;
    cubic_coeffs(r - f, w);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <int VEC>
__global__ void __launch_bounds__(TPB)
bicubic_fwd_kernel(const float* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo, float sh, float sw,
                   float* __restrict__ y) {
    const int Cv = C / VEC;
    const long total = (long)N * Ho * Wo * Cv;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % Cv) * VEC;
        long u = idx / Cv;
        const int ox = (int)(u % Wo); u /= Wo;
        const int oy = (int)(u % Ho);
        const int n = (int)(u / Ho);
        int iy, ix; float wy[4], wx[4];
        cubic_src(oy, sh, iy, wy);
        cubic_src(ox, sw, ix, wx);
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int yy = clampi(iy - 1 + a, 0, H - 1);
            float row[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) row[e] = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int xx = clampi(ix - 1 + b, 0, W - 1);
                const float* s = x + ((long)(n * H + yy) * W + xx) * C + c;
                if (VEC == 4) {
                    const float4 v = ld4(s);
                    row[0] += wx[b] * v.x; row[1 % VEC] += wx[b] * v.y; row[2 % VEC] += wx[b] * v.z; row[3 % VEC] += wx[b] * v.w;
                } else {
                    row[0] += wx[b] * __ldg(s);
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += wy[a] * row[e];
        }
        float* o = y + ((long)(n * Ho + oy) * Wo + ox) * C + c;
        if (VEC == 4) st4(o, make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]));
        else o[0] = acc[0];
    }
}

// adjoint: gx (zeroed by the caller) += W^T gy.  One thread per OUTPUT-grid element scatters its 16 weighted taps
// (fp32 atomics: the only user is the 1-channel x0.5 resize of pred_real_depth_hr, main_sr_model.py:361).
__global__ void __launch_bounds__(TPB)
bicubic_bwd_kernel(const float* __restrict__ gy, int N, int H, int W, int C, int Ho, int Wo, float sh, float sw,
                   float* __restrict__ gx) {
    const long total = (long)N * Ho * Wo * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        long u = idx / C;
        const int ox = (int)(u % Wo); u /= Wo;
        const int oy = (int)(u % Ho);
        const int n = (int)(u / Ho);
        int iy, ix; float wy[4], wx[4];
        cubic_src(oy, sh, iy, wy);
        cubic_src(ox, sw, ix, wx);
        const float g = gy[idx];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int yy = clampi(iy - 1 + a, 0, H - 1);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int xx = clampi(ix - 1 + b, 0, W - 1);
                atomicAdd(gx + ((long)(n * H + yy) * W + xx) * C + c, g * wy[a] * wx[b]);
            }
        }
    }
}

__global__ void __launch_bounds__(TPB)
nearest_fwd_kernel(const float* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo, float sh, float sw,
                   float* __restrict__ y) {
    const long total = (long)N * Ho * Wo * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        long u = idx / C;
        const int ox = (int)(u % Wo); u /= Wo;
        const int oy = (int)(u % Ho);
        const int n = (int)(u / Ho);
        const int yy = min((int)floorf((float)oy * sh), H - 1);
        const int xx = min((int)floorf((float)ox * sw), W - 1);
        y[idx] = __ldg(x + ((long)(n * H + yy) * W + xx) * C + c);
    }
}

// ------------------------------------------------------------------------------------------
// Exact x2 / x0.5 bicubic (the only factors main_sr_model.py uses: :279-293 x2 of the LR feature maps, :361 / :394-398 x0.5).
// With align_corners=False the source coordinate of output o is o/2 - 1/4 (x2: fractions 3/4, 1/4 alternate) or 2o + 1/2
// (x0.5: always 1/2), so the 4 tap weights are two fixed vectors and neighbouring outputs share their taps.  The generic
// kernel above loads 16 taps per output element from L1 (ncu: 15-18 % of the HBM peak, L1-bandwidth bound); these forms
// run the filter separably with a rolling window of horizontally filtered rows in registers: ~2 loads per output float4 (x2)
// and exactly the 16 input bytes per output (x0.5).  Same operation order as the generic kernel (horizontal taps
// ascending, then rows ascending).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 f4_fma(float w, const float4 v, const float4 a) {
    return make_float4(a.x + w * v.x, a.y + w * v.y, a.z + w * v.z, a.w + w * v.w);
}
__device__ __forceinline__ float4 f4_mul(float w, const float4 v) { return make_float4(w * v.x, w * v.y, w * v.z, w * v.w); }

#define UP2_ROWS 8             // input rows per thread (16 output rows)
// NHWC, C % 4 == 0: thread = (input column k, channel quad); it produces output columns 2k, 2k+1 of 2 * UP2_ROWS output rows
__global__ void __launch_bounds__(TPB)
bicubic_up2_kernel(const float4* __restrict__ x, int N, int H, int W, int Cv, float4* __restrict__ y) {
    const int cols_per_block = TPB / Cv;
    const int c = threadIdx.x % Cv, k = blockIdx.x * cols_per_block + threadIdx.x / Cv;
    const int m0 = blockIdx.y * UP2_ROWS, n = blockIdx.z;
    if (k >= W) return;
    float wa[4], wb[4];                       // even outputs: taps k-2..k+1, t = 3/4;  odd outputs: taps k-1..k+2, t = 1/4
    cubic_coeffs(0.75f, wa);
    cubic_coeffs(0.25f, wb);
    int xs[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) xs[b] = clampi(k - 2 + b, 0, W - 1);
    const float4* xn = x + (long)n * H * W * Cv + c;
    float4* yn = y + (long)n * 4 * H * W * Cv + c;
    const int Wo = 2 * W;
    float4 he[5], ho[5];                      // horizontally filtered input rows m-2 .. m+2 (even / odd output column)
    auto hrow = [&](int r, float4& e, float4& o) {
        const float4* row = xn + (long)clampi(r, 0, H - 1) * W * Cv;
        float4 v[5];
#pragma unroll
        for (int b = 0; b < 5; ++b) v[b] = row[(long)xs[b] * Cv];
        e = f4_mul(wa[0], v[0]); e = f4_fma(wa[1], v[1], e); e = f4_fma(wa[2], v[2], e); e = f4_fma(wa[3], v[3], e);
        o = f4_mul(wb[0], v[1]); o = f4_fma(wb[1], v[2], o); o = f4_fma(wb[2], v[3], o); o = f4_fma(wb[3], v[4], o);
    };
#pragma unroll
    for (int q = 0; q < 4; ++q) hrow(m0 - 2 + q, he[q], ho[q]);
#pragma unroll
    for (int q = 0; q < UP2_ROWS; ++q) {
        const int m = m0 + q;
        if (m >= H) break;
        hrow(m + 2, he[4], ho[4]);
        // output row 2m: input rows m-2..m+1 (weights wa); output row 2m+1: rows m-1..m+2 (weights wb)
        float4 a0 = f4_mul(wa[0], he[0]), a1 = f4_mul(wa[0], ho[0]), b0 = f4_mul(wb[0], he[1]), b1 = f4_mul(wb[0], ho[1]);
#pragma unroll
        for (int t = 1; t < 4; ++t) {
            a0 = f4_fma(wa[t], he[t], a0); a1 = f4_fma(wa[t], ho[t], a1);
            b0 = f4_fma(wb[t], he[t + 1], b0); b1 = f4_fma(wb[t], ho[t + 1], b1);
        }
        float4* o0 = yn + ((long)(2 * m) * Wo + 2 * k) * Cv;
        o0[0] = a0; o0[Cv] = a1;
        o0[(long)Wo * Cv] = b0; o0[(long)Wo * Cv + Cv] = b1;
#pragma unroll
        for (int t = 0; t < 4; ++t) { he[t] = he[t + 1]; ho[t] = ho[t + 1]; }
    }
}

#define DN2_ROWS 4             // output rows per thread
// planes (C = 1), W % 4 == 0, 16-byte aligned rows: thread = output columns ox, ox+1 (ox even) of DN2_ROWS output rows
__global__ void __launch_bounds__(TPB)
bicubic_down2_kernel(const float* __restrict__ x, int H, int W, float* __restrict__ y) {
    const int Ho = H / 2, Wo = W / 2;
    const int lane = threadIdx.x & 31;
    const int ox = (blockIdx.x * 32 + lane) * 2;
    const int oy0 = (blockIdx.y * (TPB / 32) + (threadIdx.x >> 5)) * DN2_ROWS;
    const long pl = blockIdx.z;
    const float* xp = x + pl * H * W;
    float* yp = y + pl * Ho * Wo;
    float w4[4];
    cubic_coeffs(0.5f, w4);
    const bool act = ox < Wo;
    const int xc = act ? 2 * ox : 0;          // input columns 2ox-1 .. 2ox+4 = left, the float4 at 2ox, right
    // Output row oy reads input rows 2oy-1 .. 2oy+2: the 2 * DN2_ROWS + 2 rows of the thread are loaded FIRST, all at once
    // (10 independent 16-byte loads in flight per thread; the rolling version that loaded two rows per output row ran at
    // 63 % of the HBM peak, bound by load latency - r2a), then filtered horizontally and vertically from registers.
    constexpr int NR = 2 * DN2_ROWS + 2;
    float4 v[NR];
    float sl[NR], sr[NR];
    const bool needl = xc > 0 && lane == 0, needr = xc + 4 < W && lane == 31;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const float* row = xp + (long)clampi(2 * oy0 - 1 + k, 0, H - 1) * W;
        v[k] = ld4(row + xc);
        sl[k] = needl ? __ldg(row + xc - 1) : 0.f;
        sr[k] = needr ? __ldg(row + xc + 4) : 0.f;
    }
    float h0[NR], h1[NR];                     // horizontally filtered rows: even / odd output column of the pair
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        float l = __shfl_up_sync(0xffffffffu, v[k].w, 1), rr = __shfl_down_sync(0xffffffffu, v[k].x, 1);
        if (xc == 0) l = v[k].x; else if (lane == 0) l = sl[k];
        if (xc + 4 >= W) rr = v[k].w; else if (lane == 31) rr = sr[k];
        float e = w4[0] * l; e += w4[1] * v[k].x; e += w4[2] * v[k].y; e += w4[3] * v[k].z;
        float o = w4[0] * v[k].y; o += w4[1] * v[k].z; o += w4[2] * v[k].w; o += w4[3] * rr;
        h0[k] = e; h1[k] = o;
    }
#pragma unroll
    for (int q = 0; q < DN2_ROWS; ++q) {
        const int oy = oy0 + q;
        float a = w4[0] * h0[2 * q], b = w4[0] * h1[2 * q];
#pragma unroll
        for (int t = 1; t < 4; ++t) { a += w4[t] * h0[2 * q + t]; b += w4[t] * h1[2 * q + t]; }
        if (act && oy < Ho) *reinterpret_cast<float2*>(yp + (long)oy * Wo + ox) = make_float2(a, b);
    }
}

extern "C" int dsr_bicubic_fwd(const float* x, int N, int H, int W, int C, int Ho, int Wo, float* y, void* stream) {
    DSR_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "bad arguments");
    const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
    const bool al16 = !((uintptr_t)x & 15) && !((uintptr_t)y & 15);
    if (Ho == 2 * H && Wo == 2 * W && (C & 3) == 0 && C / 4 <= TPB && TPB % (C / 4) == 0 && al16 && N <= 65535 &&
        dsr_cdiv(H, UP2_ROWS) <= 65535) {
        const int Cv = C / 4;
        bicubic_up2_kernel<<<dim3(dsr_cdiv(W, TPB / Cv), dsr_cdiv(H, UP2_ROWS), N), TPB, 0, ST(stream)>>>((const float4*)x, N, H, W, Cv, (float4*)y);
        return dsr_check_launch("bicubic_fwd (x2)");
    }
    if (C == 1 && H == 2 * Ho && W == 2 * Wo && (W & 3) == 0 && al16 && N <= 65535 && dsr_cdiv(Ho, (TPB / 32) * DN2_ROWS) <= 65535) {
        bicubic_down2_kernel<<<dim3(dsr_cdiv(Wo, 64), dsr_cdiv(Ho, (TPB / 32) * DN2_ROWS), N), TPB, 0, ST(stream)>>>(x, H, W, y);
        return dsr_check_launch("bicubic_fwd (x0.5)");
    }
    if ((C & 3) == 0 && !((uintptr_t)x & 15) && !((uintptr_t)y & 15))
        bicubic_fwd_kernel<4><<<dsr_grid((long)N * Ho * Wo * (C / 4), TPB), TPB, 0, ST(stream)>>>(x, N, H, W, C, Ho, Wo, sh, sw, y);
    else
        bicubic_fwd_kernel<1><<<dsr_grid((long)N * Ho * Wo * C, TPB), TPB, 0, ST(stream)>>>(x, N, H, W, C, Ho, Wo, sh, sw, y);
    return dsr_check_launch("bicubic_fwd");
}
extern "C" int dsr_bicubic_bwd(const float* gy, int N, int H, int W, int C, int Ho, int Wo, float* gx, void* stream) {
    DSR_REQUIRE(gy && gx && N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "bad arguments");
    const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
    bicubic_bwd_kernel<<<dsr_grid((long)N * Ho * Wo * C, TPB), TPB, 0, ST(stream)>>>(gy, N, H, W, C, Ho, Wo, sh, sw, gx);
    return dsr_check_launch("bicubic_bwd");
}
extern "C" int dsr_nearest_fwd(const float* x, int N, int H, int W, int C, int Ho, int Wo, float* y, void* stream) {
    DSR_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "bad arguments");
    const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
    nearest_fwd_kernel<<<dsr_grid((long)N * Ho * Wo * C, TPB), TPB, 0, ST(stream)>>>(x, N, H, W, C, Ho, Wo, sh, sw, y);
    return dsr_check_launch("nearest_fwd");
}

// ------------------------------------------------------------------------------------------
// On-disk formats either side of the path
//   input : uint16 PNG depth in millimetres, clipped at `max_mm` (5100), -> d / max_mm * 2 - 1 (float64 arithmetic as numpy
//           does for int / int, then float32), data/my_main_dataset.py:35-52; uint8 RGB -> (x - 127.5) / 127.5 (float32)
//   output: clip((pred + 1) / 2, 0, 1) * 5100 -> uint16 (truncation), rows [crop, H - crop), models/main_model.py:321-333
// ------------------------------------------------------------------------------------------
__global__ void u16_to_depth_kernel(const unsigned short* __restrict__ in, long n, int max_mm, float* __restrict__ out) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int v = in[i];
        const double d = (double)(v > max_mm ? max_mm : v) / (double)max_mm;
        out[i] = (float)(d * 2.0 - 1.0);
    }
}
// HWC uint8 (as decoded from the file) -> CHW float32 planes
__global__ void u8_to_image_kernel(const unsigned char* __restrict__ in, int N, int H, int W, int C, float* __restrict__ out) {
    const long plane = (long)H * W, total = (long)N * plane * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long p = i / C, n = p / plane, q = p - n * plane;
        out[(n * C + c) * plane + q] = ((float)in[i] - 127.5f) / 127.5f;
    }
}
__global__ void depth_to_u16_kernel(const float* __restrict__ pred, int N, int H, int W, int crop, float scale,
                                    unsigned short* __restrict__ out) {
    const int Ho = H - 2 * crop;
    const long total = (long)N * Ho * W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const long r = i / W;
        const int h = (int)(r % Ho), n = (int)(r / Ho);
        float v = (pred[((long)n * H + h + crop) * W + w] + 1.f) / 2.f;
        v = fminf(fmaxf(v, 0.f), 1.f) * scale;
        out[i] = (unsigned short)v;                 // numpy astype(np.uint16): truncation
    }
}
extern "C" int dsr_u16_to_depth(const unsigned short* in, long n, int max_mm, float* out, void* stream) {
    DSR_REQUIRE(in && out && n > 0 && max_mm > 0, "bad arguments");
    u16_to_depth_kernel<<<dsr_grid(n, TPB), TPB, 0, ST(stream)>>>(in, n, max_mm, out);
    return dsr_check_launch("u16_to_depth");
}
extern "C" int dsr_u8_to_image(const unsigned char* in, int N, int H, int W, int C, float* out, void* stream) {
    DSR_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0, "bad arguments");
    u8_to_image_kernel<<<dsr_grid((long)N * H * W * C, TPB), TPB, 0, ST(stream)>>>(in, N, H, W, C, out);
    return dsr_check_launch("u8_to_image");
}
extern "C" int dsr_depth_to_u16(const float* pred, int N, int H, int W, int crop, float scale, unsigned short* out, void* stream) {
    DSR_REQUIRE(pred && out && N > 0 && W > 0 && crop >= 0 && H > 2 * crop, "bad arguments");
    depth_to_u16_kernel<<<dsr_grid((long)N * (H - 2 * crop) * W, TPB), TPB, 0, ST(stream)>>>(pred, N, H, W, crop, scale, out);
    return dsr_check_launch("depth_to_u16");
}
