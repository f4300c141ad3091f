// Training-time augmentation of MyUnalignedDataset.trasform (data/my_main_dataset.py:56-90) as ONE gather pass per tensor:
// Rotate (cv2.warpAffine INTER_LINEAR / BORDER_REFLECT_101, albumentations 0.4.6 `rotate`) -> RandomCrop or
// PadIfNeeded(reflect-101) -> HorizontalFlip -> clip to [-1, 1], plus the integer-factor INTER_AREA down-scale in front of it.
// The coordinate work restates warpAffine's fixed-point scheme bit for bit (1/1024-pixel coordinates rounded half-to-even,
// reduced to 1/32 pixel, a float table of bilinear weights, products and sums rounded one by one - no FMA contraction), so the
// result equals cv2's on the same parameters.  HBM-bound: reads <= 4 taps (L1 / L2 hits between neighbours), writes 4 B.
#include "common.cuh"

#define ST(s) ((cudaStream_t)(s))

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; if (p >= n) p = 2 * (n - 1) - p; }
    return p;
}

// minv: double [N][6] = inverse affine map (dst -> src) of the rotation, ipar: int [N][4] = {rotate?, top, left, flip}
__global__ void augment_gather_kernel(const float* __restrict__ src, int N, int C, int Hs, int Ws, const double* __restrict__ minv,
                                      const int* __restrict__ ipar, float* __restrict__ dst, int Hd, int Wd) {
    const long total = (long)N * Hd * Wd;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % Wd);
        const int i = (int)((idx / Wd) % Hd);
        const int n = (int)(idx / ((long)Wd * Hd));
        const int rot = ipar[4 * n], top = ipar[4 * n + 1], left = ipar[4 * n + 2], flip = ipar[4 * n + 3];
        // position in the rotated (load-size) frame: flip acts on the cropped / padded image, reflect-101 maps padding back
        const int x = reflect101((flip ? Wd - 1 - j : j) + left, Ws);
        const int y = reflect101(i + top, Hs);
        const float* s = src + (long)n * C * Hs * Ws;
        float* d = dst + (long)n * C * Hd * Wd + (long)i * Wd + j;
        if (!rot) {
            for (int c = 0; c < C; ++c) d[(long)c * Hd * Wd] = fminf(fmaxf(s[((long)c * Hs + y) * Ws + x], -1.f), 1.f);
            continue;
        }
        const double* m = minv + 6 * n;
        // imgwarp.cpp WarpAffineInvoker: AB_BITS = 10, INTER_BITS = 5, round_delta = 16; saturate_cast<int>(double) = rint
        const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0));
        const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0));
        const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
        const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
        const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
        const int sx = X >> 5, sy = Y >> 5;
        const float fx = __fmul_rn((float)(X & 31), 0.03125f), fy = __fmul_rn((float)(Y & 31), 0.03125f);
        const float vx0 = __fsub_rn(1.f, fx), vy0 = __fsub_rn(1.f, fy);
        const float w0 = __fmul_rn(vy0, vx0), w1 = __fmul_rn(vy0, fx), w2 = __fmul_rn(fy, vx0), w3 = __fmul_rn(fy, fx);
        const int x0 = reflect101(sx, Ws), x1 = reflect101(sx + 1, Ws), y0 = reflect101(sy, Hs), y1 = reflect101(sy + 1, Hs);
        for (int c = 0; c < C; ++c) {
            const float* p = s + (long)c * Hs * Ws;
            float v = __fmul_rn(p[(long)y0 * Ws + x0], w0);
            v = __fadd_rn(v, __fmul_rn(p[(long)y0 * Ws + x1], w1));
            v = __fadd_rn(v, __fmul_rn(p[(long)y1 * Ws + x0], w2));
            v = __fadd_rn(v, __fmul_rn(p[(long)y1 * Ws + x1], w3));
            d[(long)c * Hd * Wd] = fminf(fmaxf(v, -1.f), 1.f);
        }
    }
}

// cv2.resize(INTER_AREA) for integer down-scale factors (resizeAreaFast): float running sum over the box, rows first, * 1/area
__global__ void resize_area_int_kernel(const float* __restrict__ src, long planes, int H, int W, int fy, int fx, float* __restrict__ dst) {
    const int Ho = H / fy, Wo = W / fx;
    const float scale = 1.f / (float)(fy * fx);
    const long total = planes * Ho * Wo;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % Wo);
        const int i = (int)((idx / Wo) % Ho);
        const long pl = idx / ((long)Wo * Ho);
        const float* s = src + pl * H * W;
        float acc = 0.f;
        for (int dy = 0; dy < fy; ++dy)
            for (int dx = 0; dx < fx; ++dx) acc = __fadd_rn(acc, s[(long)(i * fy + dy) * W + j * fx + dx]);
        dst[idx] = __fmul_rn(acc, scale);
    }
}

extern "C" int dsr_augment_gather(const float* src, int N, int C, int Hs, int Ws, const double* minv, const int* ipar, float* dst,
                                  int Hd, int Wd, void* stream) {
    DSR_REQUIRE(src && minv && ipar && dst && N > 0 && C > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0, "bad arguments");
    augment_gather_kernel<<<dsr_grid((long)N * Hd * Wd, 256), 256, 0, ST(stream)>>>(src, N, C, Hs, Ws, minv, ipar, dst, Hd, Wd);
    return dsr_check_launch("augment_gather");
}
extern "C" int dsr_resize_area_int(const float* src, long planes, int H, int W, int fy, int fx, float* dst, void* stream) {
    DSR_REQUIRE(src && dst && planes > 0 && fy > 0 && fx > 0 && H % fy == 0 && W % fx == 0, "integer down-scale factors only");
    resize_area_int_kernel<<<dsr_grid(planes * (H / fy) * (W / fx), 256), 256, 0, ST(stream)>>>(src, planes, H, W, fy, fx, dst);
    return dsr_check_launch("resize_area_int");
}
