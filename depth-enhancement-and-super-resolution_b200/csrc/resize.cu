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

extern "C" int dsr_bicubic_fwd(const float* x, int N, int H, int W, int C, int Ho, int Wo, float* y, void* stream) {
    DSR_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "bad arguments");
    const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
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
