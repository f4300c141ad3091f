// Kernels that only the translation block needs (models/translation_network.py of the reference):
//   * SurfaceNormals (:329-360): field-of-view normals - P = depth * grid(482 x 642, 60 deg), reflection pad, half differences
//     (with the reference's one-pixel up / left shift of the taps), n = -(gx x gy) / max(|.|, 1e-8); forward + backward
//   * CosSimLoss (:310-316): mean(1 - cos(x, y)) over pixels, channel dimension 1; forward + gradient w.r.t. x
// NCHW fp32 planes.  Off the north-star path: straightforward one-thread-per-pixel kernels.
#include "common.cuh"
#include "../../include/dsr_b200.h"

#define ST(s) ((cudaStream_t)(s))
#define TT 256

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }
// grid(i, j) of generate_grid(482, 642, 60) narrowed to the h x w window (:338-350); tan(30 deg) passed in as `t`
__device__ __forceinline__ void fov_grid(int i, int j, int h, int w, float t, float& gx, float& gy) {
    const int ph = (482 - h) / 2, pw = (642 - w) / 2;
    const float xi = (float)(pw + 1 + j + 1), yi = (float)(ph + 1 + i + 1);        // torch.arange(1, n + 1)[narrow offset + index]
    gx = (xi - (642.f + 1.f) / 2.f) / (642.f / 2.f) * t;
    gy = -(yi - (482.f + 1.f) / 2.f) / (482.f / 2.f) * t * (482.f / 642.f);
}
__device__ __forceinline__ void fov_point(const float* __restrict__ d, int h, int w, int i, int j, float t, float P[3]) {
    i = reflect1(i, h); j = reflect1(j, w);
    float gx, gy;
    fov_grid(i, j, h, w, t, gx, gy);
    const float v = d[(long)i * w + j];
    P[0] = v * gx; P[1] = v * gy; P[2] = v;
}
// gx(i,j) = (P(i-1,j-1) - P(i-1,j+1)) / 2,  gy(i,j) = (P(i+1,j-1) - P(i-1,j-1)) / 2   (:352-353, padded-coordinate narrows)
__device__ __forceinline__ void fov_grads(const float* __restrict__ d, int h, int w, int i, int j, float t, float gx[3], float gy[3]) {
    float A[3], B[3], C[3];
    fov_point(d, h, w, i - 1, j - 1, t, A);
    fov_point(d, h, w, i - 1, j + 1, t, B);
    fov_point(d, h, w, i + 1, j - 1, t, C);
#pragma unroll
    for (int k = 0; k < 3; ++k) { gx[k] = A[k] / 2.f - B[k] / 2.f; gy[k] = C[k] / 2.f - A[k] / 2.f; }
}
__global__ void fov_normals_fwd_kernel(const float* __restrict__ d, int B, int h, int w, float t, float* __restrict__ out) {
    const long plane = (long)h * w, total = (long)B * plane;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / plane);
        const long p = idx - (long)b * plane;
        const int i = (int)(p / w), j = (int)(p - (long)i * w);
        float gx[3], gy[3];
        fov_grads(d + b * plane, h, w, i, j, t, gx, gy);
        const float c0 = gx[1] * gy[2] - gx[2] * gy[1], c1 = gx[2] * gy[0] - gx[0] * gy[2], c2 = gx[0] * gy[1] - gx[1] * gy[0];
        const float nrm = sqrtf(c0 * c0 + c1 * c1 + c2 * c2), den = nrm > 1e-8f ? nrm : 1e-8f;
        float* o = out + (long)b * 3 * plane + p;
        o[0] = -c0 / den; o[plane] = -c1 / den; o[2 * plane] = -c2 / den;
    }
}
// adjoint, scattered: gd (zeroed by the caller) += dL/dd
__global__ void fov_normals_bwd_kernel(const float* __restrict__ d, const float* __restrict__ g, int B, int h, int w, float t,
                                       float* __restrict__ gd) {
    const long plane = (long)h * w, total = (long)B * plane;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / plane);
        const long p = idx - (long)b * plane;
        const int i = (int)(p / w), j = (int)(p - (long)i * w);
        const float* db = d + b * plane;
        float gx[3], gy[3];
        fov_grads(db, h, w, i, j, t, gx, gy);
        const float c[3] = {gx[1] * gy[2] - gx[2] * gy[1], gx[2] * gy[0] - gx[0] * gy[2], gx[0] * gy[1] - gx[1] * gy[0]};
        const float nrm = sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        const float* gb = g + (long)b * 3 * plane + p;
        const float gn[3] = {gb[0], gb[plane], gb[2 * plane]};
        float dc[3];                                      // dL/dcrs with n = -crs / max(|crs|, 1e-8)
        if (nrm > 1e-8f) {
            const float u[3] = {c[0] / nrm, c[1] / nrm, c[2] / nrm};
            const float dot = u[0] * gn[0] + u[1] * gn[1] + u[2] * gn[2];
#pragma unroll
            for (int k = 0; k < 3; ++k) dc[k] = -(gn[k] - u[k] * dot) / nrm;
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) dc[k] = -gn[k] / 1e-8f;
        }
        // crs = gx x gy  =>  dL/dgx = gy x dc,  dL/dgy = dc x gx
        const float dgx[3] = {gy[1] * dc[2] - gy[2] * dc[1], gy[2] * dc[0] - gy[0] * dc[2], gy[0] * dc[1] - gy[1] * dc[0]};
        const float dgy[3] = {dc[1] * gx[2] - dc[2] * gx[1], dc[2] * gx[0] - dc[0] * gx[2], dc[0] * gx[1] - dc[1] * gx[0]};
        // gx = A/2 - B/2, gy = C/2 - A/2 with A = P(i-1,j-1), B = P(i-1,j+1), C = P(i+1,j-1); P(q) = d(q) * grid(q)
        const int qi[3] = {reflect1(i - 1, h), reflect1(i - 1, h), reflect1(i + 1, h)};
        const int qj[3] = {reflect1(j - 1, w), reflect1(j + 1, w), reflect1(j - 1, w)};
        float dP[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { dP[0][k] = (dgx[k] - dgy[k]) / 2.f; dP[1][k] = -dgx[k] / 2.f; dP[2][k] = dgy[k] / 2.f; }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float grx, gry;
            fov_grid(qi[q], qj[q], h, w, t, grx, gry);
            atomicAdd(gd + b * plane + (long)qi[q] * w + qj[q], dP[q][0] * grx + dP[q][1] * gry + dP[q][2]);
        }
    }
}

// cosine similarity along the channel dimension (nn.CosineSimilarity(dim=1), eps 1e-8): x.y / (max(|x|, eps) * max(|y|, eps))
__global__ void cos_sim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ mask,
                                   int B, int C, long plane, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long)B * plane; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / plane, p = idx - b * plane;
        float dot = 0.f, nx = 0.f, ny = 0.f;
        for (int c = 0; c < C; ++c) {
            const float a = x[(b * C + c) * plane + p], v = y[(b * C + c) * plane + p];
            dot += a * v; nx += a * a; ny += v * v;
        }
        const float l = 1.f - dot / (fmaxf(sqrtf(nx), 1e-8f) * fmaxf(sqrtf(ny), 1e-8f));
        acc += (double)(mask ? l * mask[idx] : l);
    }
    acc = block_sum<double>(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}
// gx = coef * (*gscale) * d sum(1 - cos) / dx
__global__ void cos_sim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ mask,
                                   int B, int C, long plane, const float* __restrict__ gscale, float coef, float* __restrict__ gx) {
    const float gs0 = coef * (gscale ? *gscale : 1.f);
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long)B * plane; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / plane, p = idx - b * plane;
        const float gs = mask ? gs0 * mask[idx] : gs0;
        float dot = 0.f, nx = 0.f, ny = 0.f;
        for (int c = 0; c < C; ++c) {
            const float a = x[(b * C + c) * plane + p], v = y[(b * C + c) * plane + p];
            dot += a * v; nx += a * a; ny += v * v;
        }
        const float rx = sqrtf(nx), ry = sqrtf(ny), mx = fmaxf(rx, 1e-8f), my = fmaxf(ry, 1e-8f);
        for (int c = 0; c < C; ++c) {
            const float a = x[(b * C + c) * plane + p], v = y[(b * C + c) * plane + p];
            // d cos / dx = y / (mx my) - dot * (d mx / dx) / (mx^2 my),  d mx / dx = x / rx when rx > eps, else 0
            float dcos = v / (mx * my);
            if (rx > 1e-8f) dcos -= dot * a / (rx * mx * mx * my);
            gx[(b * C + c) * plane + p] = -gs * dcos;
        }
    }
}

extern "C" int dsr_fov_normals_fwd(const float* depth, int B, int H, int W, float* out, void* stream) {
    DSR_REQUIRE(depth && out && B > 0 && H >= 2 && W >= 2 && H <= 480 && W <= 640, "bad arguments (the 482 x 642 grid holds at most 480 x 640)");
    fov_normals_fwd_kernel<<<dsr_grid((long)B * H * W, TT), TT, 0, ST(stream)>>>(depth, B, H, W, tanf(60.f / 2.f / 180.f * 3.14159265358979323846f), out);
    return dsr_check_launch("fov_normals_fwd");
}
extern "C" int dsr_fov_normals_bwd(const float* depth, const float* gout, int B, int H, int W, float* gdepth, void* stream) {
    DSR_REQUIRE(depth && gout && gdepth && B > 0 && H >= 2 && W >= 2 && H <= 480 && W <= 640, "bad arguments");
    fov_normals_bwd_kernel<<<dsr_grid((long)B * H * W, TT), TT, 0, ST(stream)>>>(depth, gout, B, H, W, tanf(60.f / 2.f / 180.f * 3.14159265358979323846f), gdepth);
    return dsr_check_launch("fov_normals_bwd");
}
extern "C" int dsr_cos_sim_fwd(const float* x, const float* y, int B, int C, long plane, double* out_sum, void* stream) {
    DSR_REQUIRE(x && y && out_sum && B > 0 && C > 0 && plane > 0, "bad arguments");
    cos_sim_fwd_kernel<<<dsr_grid((long)B * plane, TT), TT, 0, ST(stream)>>>(x, y, nullptr, B, C, plane, out_sum);
    return dsr_check_launch("cos_sim_fwd");
}
extern "C" int dsr_cos_sim_bwd(const float* x, const float* y, int B, int C, long plane, const float* gscale, float coef, float* gx,
                               void* stream) {
    DSR_REQUIRE(x && y && gx && B > 0 && C > 0 && plane > 0, "bad arguments");
    cos_sim_bwd_kernel<<<dsr_grid((long)B * plane, TT), TT, 0, ST(stream)>>>(x, y, nullptr, B, C, plane, gscale, coef, gx);
    return dsr_check_launch("cos_sim_bwd");
}
extern "C" int dsr_cos_sim_masked_fwd(const float* x, const float* y, const float* mask, int B, int C, long plane, double* out_sum,
                                      void* stream) {
    DSR_REQUIRE(x && y && mask && out_sum && B > 0 && C > 0 && plane > 0, "bad arguments");
    cos_sim_fwd_kernel<<<dsr_grid((long)B * plane, TT), TT, 0, ST(stream)>>>(x, y, mask, B, C, plane, out_sum);
    return dsr_check_launch("cos_sim_masked_fwd");
}
extern "C" int dsr_cos_sim_masked_bwd(const float* x, const float* y, const float* mask, int B, int C, long plane, const float* gscale,
                                      float coef, float* gx, void* stream) {
    DSR_REQUIRE(x && y && mask && gx && B > 0 && C > 0 && plane > 0, "bad arguments");
    cos_sim_bwd_kernel<<<dsr_grid((long)B * plane, TT), TT, 0, ST(stream)>>>(x, y, mask, B, C, plane, gscale, coef, gx);
    return dsr_check_launch("cos_sim_masked_bwd");
}

// MaskedMeanDif (models/translation_network.py:288-293): sums[b] = (sum (y - x) * mask, sum mask) over sample b's plane
__global__ void masked_mean_dif_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ mask,
                                           long plane, double* __restrict__ sums) {
    __shared__ double red[32];
    const int b = blockIdx.y;
    double sd = 0.0, sm = 0.0;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += (long)gridDim.x * blockDim.x) {
        const float m = mask[b * plane + p];
        sd += (double)((y[b * plane + p] - x[b * plane + p]) * m);
        sm += (double)m;
    }
    sd = block_sum<double>(sd, red);
    sm = block_sum<double>(sm, red);
    if (threadIdx.x == 0) { atomicAdd(sums + 2 * b, sd); atomicAdd(sums + 2 * b + 1, sm); }
}
// *loss = mean_b |sums[b][0] / (sums[b][1] + 1e-6)|   (one thread)
__global__ void masked_mean_dif_finish_kernel(const double* __restrict__ sums, int B, float* __restrict__ loss) {
    double acc = 0.0;
    for (int b = 0; b < B; ++b) acc += fabs(sums[2 * b] / (sums[2 * b + 1] + 1e-6));
    *loss = (float)(acc / B);
}
// gx = -(*gscale) * sign(mean_b) * mask / (sum mask_b + 1e-6) / B
__global__ void masked_mean_dif_bwd_kernel(const float* __restrict__ mask, const double* __restrict__ sums, int B, long plane,
                                           const float* __restrict__ gscale, float* __restrict__ gx) {
    const int b = blockIdx.y;
    const double q = sums[2 * b] / (sums[2 * b + 1] + 1e-6);
    const float c = -(*gscale) * (q > 0.0 ? 1.f : (q < 0.0 ? -1.f : 0.f)) / (float)(sums[2 * b + 1] + 1e-6) / (float)B;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += (long)gridDim.x * blockDim.x)
        gx[b * plane + p] = c * mask[b * plane + p];
}
extern "C" int dsr_masked_mean_dif_fwd(const float* x, const float* y, const float* mask, int B, long plane, double* sums, float* loss,
                                       void* stream) {
    DSR_REQUIRE(x && y && mask && sums && loss && B > 0 && plane > 0, "bad arguments");
    dim3 grid((unsigned)min((plane + TT - 1) / TT, 64L), B);
    masked_mean_dif_fwd_kernel<<<grid, TT, 0, ST(stream)>>>(x, y, mask, plane, sums);
    masked_mean_dif_finish_kernel<<<1, 1, 0, ST(stream)>>>(sums, B, loss);
    return dsr_check_launch("masked_mean_dif_fwd");
}
extern "C" int dsr_masked_mean_dif_bwd(const float* mask, const double* sums, int B, long plane, const float* gscale, float* gx,
                                       void* stream) {
    DSR_REQUIRE(mask && sums && gscale && gx && B > 0 && plane > 0, "bad arguments");
    dim3 grid((unsigned)min((plane + TT - 1) / TT, 64L), B);
    masked_mean_dif_bwd_kernel<<<grid, TT, 0, ST(stream)>>>(mask, sums, B, plane, gscale, gx);
    return dsr_check_launch("masked_mean_dif_bwd");
}
