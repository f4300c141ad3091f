// Batch evaluator of the reference's depth metrics (new_metrics.py, /root/reference): per image, in one pass over the
// predicted / target / input depth maps (millimetres), the masked sums behind MAE / RMSE / PSNR (:115-131, :181-188), their
// "holes" (_h) and "no holes" (_d) variants (:133-160), the normal-based MSE_v (:17-66, :162-176: first-order point-cloud
// normals, plus-shaped dilation of the target hole map) and the 'valid'-mode 11x11 Gaussian SSIM (:81-113, :190-191).
// Everything accumulates in fp64 as the reference computes in float64; the host divides (dsr_b200/metrics.py).
#include "common.cuh"
#include "../../include/dsr_b200.h"

#define ST(s) ((cudaStream_t)(s))
#define MT 256
#define DSR_METRIC_SLOTS 16
// out[b][..]: 0 n_t, 1 sum|d|_t, 2 sum d^2_t | 3 n_h, 4 sum|d|_h, 5 sum d^2_h | 6 n_d, 7 sum|d|_d, 8 sum d^2_d |
//             9 n_v (elements), 10 sum (dn)^2_v | 11 n_ssim, 12 sum ssim

// imageio.imread(...).astype(np.float64).clip(0, max_depth) (:207-208)
__device__ __forceinline__ double clipd(float v, double mx) { const double d = (double)v; return d < 0.0 ? 0.0 : (d > mx ? mx : d); }

__device__ __forceinline__ void metric_point(const float* __restrict__ z, const double* __restrict__ ki, int H, int W, int i, int j,
                                             double mx, double P[3]) {
    // depth_to_absolute_coordinates (:49-66), 'orthogonal': K^-1 [u, v, 1] / (.)_z * depth, u = j + 0.5, v = i + 0.5
    const double u = (double)j + 0.5, v = (double)i + 0.5;
    const double a = ki[0] * u + ki[1] * v + ki[2], b = ki[3] * u + ki[4] * v + ki[5], c = ki[6] * u + ki[7] * v + ki[8];
    const double d = clipd(z[(long)i * W + j], mx);
    P[0] = a / c * d; P[1] = b / c * d; P[2] = c / c * d;
}
__device__ __forceinline__ void metric_normal(const float* __restrict__ z, const double* __restrict__ ki, int H, int W, int i, int j,
                                              double mx, double n[3]) {
    // coords_to_normals (:17-47): forward differences, the last column / row replicate their neighbour's difference
    const int ju = j < W - 1 ? j : W - 2, iv = i < H - 1 ? i : H - 2;
    double A[3], Bu[3], C[3], Dv[3];
    metric_point(z, ki, H, W, i, ju, mx, A); metric_point(z, ki, H, W, i, ju + 1, mx, Bu);
    metric_point(z, ki, H, W, iv, j, mx, C); metric_point(z, ki, H, W, iv + 1, j, mx, Dv);
    const double dxdu = Bu[0] - A[0], dydu = Bu[1] - A[1], dzdu = Bu[2] - A[2];
    const double dxdv = Dv[0] - C[0], dydv = Dv[1] - C[1], dzdv = Dv[2] - C[2];
    const double nx = dydv * dzdu - dydu * dzdv, ny = dzdv * dxdu - dzdu * dxdv, nz = dxdv * dydu - dxdu * dydv;
    const double r = sqrt(nx * nx + ny * ny + nz * nz), den = r > 1e-12 ? r : 1e-12;
    n[0] = nx / den; n[1] = ny / den; n[2] = nz / den;
}

__global__ void __launch_bounds__(MT)
metric_sums_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ input,
                   const double* __restrict__ kinv, int H, int W, float thr, double mx, double* __restrict__ out) {
    __shared__ double red[32];
    const int b = blockIdx.y;
    const long plane = (long)H * W;
    const float* p = pred + b * plane;
    const float* t = target + b * plane;
    const float* x = input + b * plane;
    const double* ki = kinv ? kinv + b * 9 : nullptr;
    double acc[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) acc[k] = 0.0;
    for (long q = (long)blockIdx.x * MT + threadIdx.x; q < plane; q += (long)gridDim.x * MT) {
        const int i = (int)(q / W), j = (int)(q - (long)i * W);
        const bool th = clipd(t[q], mx) < thr, ih = x[q] < thr;       // target_hole_map, hole_map (:224-225)
        const double d = clipd(p[q], mx) - clipd(t[q], mx), ad = fabs(d), d2 = d * d;
        if (!th) { acc[0] += 1.0; acc[1] += ad; acc[2] += d2; }
        if (!th && ih) { acc[3] += 1.0; acc[4] += ad; acc[5] += d2; }
        if (!th && !ih) { acc[6] += 1.0; acc[7] += ad; acc[8] += d2; }
        if (ki) {
            bool m = th;                                              // plus-shaped dilation of the target holes (:169-173)
            if (j > 0) m |= t[q - 1] < thr;
            if (j < W - 1) m |= t[q + 1] < thr;
            if (i > 0) m |= t[q - W] < thr;
            if (i < H - 1) m |= t[q + W] < thr;
            if (!m) {
                double np_[3], nt[3];
                metric_normal(p, ki, H, W, i, j, mx, np_);
                metric_normal(t, ki, H, W, i, j, mx, nt);
                acc[9] += 3.0;
                acc[10] += (np_[0] - nt[0]) * (np_[0] - nt[0]) + (np_[1] - nt[1]) * (np_[1] - nt[1]) + (np_[2] - nt[2]) * (np_[2] - nt[2]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        const double s = block_sum<double>(acc[k], red);
        if (threadIdx.x == 0 && s != 0.0) atomicAdd(out + b * DSR_METRIC_SLOTS + k, s);
    }
}

// SSIM of (~target_hole * pred / max_depth, ~target_hole * target / max_depth), 11x11 Gaussian sigma 1.5, 'valid' mode,
// L = 1 (:81-113): one output pixel per thread straight from global memory in fp64 (an evaluator, not a hot kernel)
__constant__ double c_win[121];
__global__ void __launch_bounds__(MT)
metric_ssim_kernel(const float* __restrict__ pred, const float* __restrict__ target, int H, int W, float thr, double mx,
                   double* __restrict__ out) {
    __shared__ double red[32];
    const int b = blockIdx.y, Ho = H - 10, Wo = W - 10;
    const long plane = (long)H * W;
    const float* p = pred + b * plane;
    const float* t = target + b * plane;
    const double inv_max = 1.0 / mx;
    double acc = 0.0, cnt = 0.0;
    for (long q = (long)blockIdx.x * MT + threadIdx.x; q < (long)Ho * Wo; q += (long)gridDim.x * MT) {
        const int i = (int)(q / Wo), j = (int)(q - (long)i * Wo);
        double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
        for (int r = 0; r < 11; ++r)
            for (int c = 0; c < 11; ++c) {
                const long o = (long)(i + r) * W + j + c;
                const bool keep = !(t[o] < thr);
                const double a = keep ? clipd(p[o], mx) * inv_max : 0.0, bb = keep ? clipd(t[o], mx) * inv_max : 0.0;
                const double w = c_win[r * 11 + c];
                m1 += w * a; m2 += w * bb; s11 += w * a * a; s22 += w * bb * bb; s12 += w * a * bb;
            }
        const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;
        const double v1 = s11 - m1 * m1, v2 = s22 - m2 * m2, v12 = s12 - m1 * m2;
        acc += ((2 * m1 * m2 + C1) * (2 * v12 + C2)) / ((m1 * m1 + m2 * m2 + C1) * (v1 + v2 + C2));
        cnt += 1.0;
    }
    acc = block_sum<double>(acc, red);
    cnt = block_sum<double>(cnt, red);
    if (threadIdx.x == 0 && cnt != 0.0) { atomicAdd(out + b * DSR_METRIC_SLOTS + 11, cnt); atomicAdd(out + b * DSR_METRIC_SLOTS + 12, acc); }
}

extern "C" int dsr_eval_metric_sums(const float* pred, const float* target, const float* input, const double* kinv, int B, int H,
                                    int W, float hole_threshold, double max_depth, int with_ssim, double* out, void* stream) {
    DSR_REQUIRE(pred && target && input && out && B > 0 && B <= 65535 && H >= 2 && W >= 2 && max_depth > 0, "bad arguments");
    DSR_REQUIRE(!with_ssim || (H >= 11 && W >= 11), "SSIM needs at least an 11 x 11 image");
    if (cudaMemsetAsync(out, 0, (size_t)B * DSR_METRIC_SLOTS * sizeof(double), ST(stream)) != cudaSuccess) {
        dsr_set_error("eval_metric_sums: memset failed"); return DSR_ERR_CUDA;
    }
    const long plane = (long)H * W;
    int gx = dsr_cdiv(plane, MT);
    const int cap = dsr_num_sms() * 8 / B + 1;
    if (gx > cap) gx = cap;
    metric_sums_kernel<<<dim3(gx, B), MT, 0, ST(stream)>>>(pred, target, input, kinv, H, W, hole_threshold, max_depth, out);
    int rc = dsr_check_launch("eval_metric_sums");
    if (rc || !with_ssim) return rc;
    static bool init = false;
    if (!init) {       // fspecial_gauss(11, 1.5) (:68-79): exp(-(x^2 + y^2) / (2 sigma^2)) on [-5, 5]^2, normalised
        double w[121], s = 0.0;
        for (int y = -5; y <= 5; ++y)
            for (int x = -5; x <= 5; ++x) { w[(y + 5) * 11 + x + 5] = exp(-((double)(x * x + y * y)) / (2.0 * 1.5 * 1.5)); s += w[(y + 5) * 11 + x + 5]; }
        for (int k = 0; k < 121; ++k) w[k] /= s;
        if (cudaMemcpyToSymbol(c_win, w, sizeof(w)) != cudaSuccess) { dsr_set_error("eval_metric_sums: constant upload failed"); return DSR_ERR_CUDA; }
        init = true;
    }
    int gs = dsr_cdiv((long)(H - 10) * (W - 10), MT);
    if (gs > cap) gs = cap;
    metric_ssim_kernel<<<dim3(gs, B), MT, 0, ST(stream)>>>(pred, target, H, W, hole_threshold, max_depth, out);
    return dsr_check_launch("eval_metric_sums (ssim)");
}
