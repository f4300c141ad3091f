// Depth-derived stencil / reduction kernels of the loss stack (HBM-bound; NCHW fp32 planes exactly
// as the reference holds them).  Reference call sites: models/main_model.py:208-230 (masks),
// :257-298 (rectangle holes), :340-417 (loss stack), models/norms.py (normals),
// models/pytorch_ssim/__init__.py (SSIM).  C-ABI entry points at the bottom (include/dsr_b200.h).
#include "common.cuh"
#include "stencil_math.cuh"
#include "../../include/dsr_b200.h"

#define TPB 256

// ------------------------------------------------------------------------------------------
// hole / valid masks  (main_model.py:208-230): hole = d <= border; valid = !dilate3x3(hole)
// ------------------------------------------------------------------------------------------
__global__ void hole_valid_kernel(const float* __restrict__ d, int B, int H, int W, float border,
                                  float* __restrict__ hole, float* __restrict__ valid) {
    long total = (long)B * H * W;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int j = (int)(idx % W);
        int i = (int)((idx / W) % H);
        const float* p = d + (idx - (long)i * W - j);
        bool any = false;
#pragma unroll
        for (int di = -1; di <= 1; ++di) {
            int ii = i + di;
            if (ii < 0 || ii >= H) continue;
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                int jj = j + dj;
                if (jj < 0 || jj >= W) continue;
                any |= (__ldg(p + (long)ii * W + jj) <= border);
            }
        }
        if (hole) hole[idx] = (d[idx] <= border) ? 1.f : 0.f;
        valid[idx] = any ? 0.f : 1.f;
    }
}

// ------------------------------------------------------------------------------------------
// rectangle holes (main_model.py:257-298 + :354-357 / :396)
//   gt = !(valid > 0.05 && covered);  masked = gt ? depth : -1;
//   extra = (masked < extra_border) || !gt      (extra_border = -inf for the real domain)
// rects: int32 [B][max_rects][4] = x, y, size_x, size_y ; counts int32 [B]
// ------------------------------------------------------------------------------------------
__global__ void rect_holes_kernel(const float* __restrict__ valid, const float* __restrict__ depth,
                                  const int* __restrict__ rects, const int* __restrict__ counts, int max_rects,
                                  int H, int W, float extra_border, unsigned char* __restrict__ gt,
                                  float* __restrict__ masked, float* __restrict__ extra) {
    extern __shared__ int srect[];
    int b = blockIdx.y;
    int n = counts[b];
    for (int t = threadIdx.x; t < n * 4; t += blockDim.x) srect[t] = rects[(long)b * max_rects * 4 + t];
    __syncthreads();
    long plane = (long)H * W;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += (long)gridDim.x * blockDim.x) {
        int x = (int)(p % W), y = (int)(p / W);
        bool cov = false;
        for (int r = 0; r < n; ++r) {
            int rx = srect[4 * r], ry = srect[4 * r + 1], sx = srect[4 * r + 2], sy = srect[4 * r + 3];
            cov |= (x >= rx) & (x < rx + sx) & (y >= ry) & (y < ry + sy);
        }
        long o = (long)b * plane + p;
        bool g = !((valid[o] > 0.05f) && cov);
        float m = g ? depth[o] : -1.f;
        gt[o] = g ? 1 : 0;
        masked[o] = m;
        if (extra) extra[o] = ((m < extra_border) || !g) ? 1.f : 0.f;
    }
}

// ------------------------------------------------------------------------------------------
// normals
// ------------------------------------------------------------------------------------------
__global__ void normals_old_fwd_kernel(const float* __restrict__ d, int B, int H, int W, float scale,
                                       float* __restrict__ out) {
    long plane = (long)H * W, total = (long)B * plane;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int b = (int)(idx / plane);
        long p = idx - (long)b * plane;
        int i = (int)(p / W), j = (int)(p % W);
        float n[3];
        old_normal_fwd(d + (long)b * plane, H, W, i, j, scale, n);
        float* o = out + (long)b * 3 * plane + p;
        o[0] = n[0]; o[plane] = n[1]; o[2 * plane] = n[2];
    }
}
__global__ void normals_old_bwd_kernel(const float* __restrict__ d, const float* __restrict__ g, int B, int H,
                                       int W, float scale, float* __restrict__ gd) {
    long plane = (long)H * W, total = (long)B * plane;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int b = (int)(idx / plane);
        long p = idx - (long)b * plane;
        int i = (int)(p / W), j = (int)(p % W);
        gd[idx] = old_normal_bwd(d + (long)b * plane, g + (long)b * 3 * plane, plane, H, W, i, j, scale);
    }
}
__global__ void normals_new_fwd_kernel(const float* __restrict__ d, const double* __restrict__ cams, int B,
                                       int H, int W, float* __restrict__ out) {
    long plane = (long)H * W, total = (long)B * plane;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int b = (int)(idx / plane);
        long p = idx - (long)b * plane;
        int i = (int)(p / W), j = (int)(p % W);
        float n[3];
        new_normal_fwd(d + (long)b * plane, cams + b * DSR_CAM_DOUBLES, H, W, i, j, n);
        float* o = out + (long)b * 3 * plane + p;
        o[0] = n[0]; o[plane] = n[1]; o[2 * plane] = n[2];
    }
}
__global__ void normals_new_bwd_kernel(const float* __restrict__ d, const float* __restrict__ g,
                                       const double* __restrict__ cams, int B, int H, int W,
                                       float* __restrict__ gd) {
    long plane = (long)H * W, total = (long)B * plane;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int b = (int)(idx / plane);
        long p = idx - (long)b * plane;
        int i = (int)(p / W), j = (int)(p % W);
        gd[idx] = new_normal_bwd(d + (long)b * plane, g + (long)b * 3 * plane, plane,
                                 cams + b * DSR_CAM_DOUBLES, H, W, i, j);
    }
}

// ------------------------------------------------------------------------------------------
// total variation (main_model.py:15-19): sum of squared forward differences, x is (BC, H, W)
// ------------------------------------------------------------------------------------------
__global__ void tv_fwd_kernel(const float* __restrict__ x, long planes, int H, int W, double* __restrict__ out) {
    __shared__ double red[32];
    long plane = (long)H * W, total = planes * plane;
    double acc = 0.0;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long p = idx % plane;
        int i = (int)(p / W), j = (int)(p % W);
        float v = x[idx];
        float a = 0.f;
        if (j < W - 1) { float t = v - x[idx + 1]; a += t * t; }
        if (i < H - 1) { float t = v - x[idx + W]; a += t * t; }
        acc += (double)a;
    }
    acc = block_sum<double>(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}
__global__ void tv_bwd_kernel(const float* __restrict__ x, long planes, int H, int W, const float* __restrict__ gscale,
                              float coef, float* __restrict__ gx) {
    long plane = (long)H * W, total = planes * plane;
    float gs = coef * (gscale ? *gscale : 1.f) * 2.f;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long p = idx % plane;
        int i = (int)(p / W), j = (int)(p % W);
        float v = x[idx];
        float a = 0.f;
        if (j < W - 1) a += v - x[idx + 1];
        if (j > 0) a -= x[idx - 1] - v;
        if (i < H - 1) a += v - x[idx + W];
        if (i > 0) a -= x[idx - W] - v;
        gx[idx] = gs * a;
    }
}

// ------------------------------------------------------------------------------------------
// masked L1 / L2 sums:  t = (a*m1)*m2 - (b*m1)*m2 ; out[0] += sum|t| ; out[1] += sum t^2
// a, b are (B, C, H, W); masks are (B, 1, H, W) (m2 may be null).  main_model.py:352,371-372,383-398
// ------------------------------------------------------------------------------------------
__global__ void masked_diff_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                       const float* __restrict__ m1, const float* __restrict__ m2, int B, int C,
                                       long plane, double* __restrict__ out) {
    __shared__ double red[32];
    long total = (long)B * C * plane;
    double s1 = 0.0, s2 = 0.0;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long bc = idx / plane;
        long mo = (bc / C) * plane + (idx - bc * plane);
        float m = m1[mo];
        float ta = a[idx] * m, tb = b[idx] * m;
        if (m2) { float mm = m2[mo]; ta *= mm; tb *= mm; }
        float t = ta - tb;
        s1 += (double)fabsf(t);
        s2 += (double)(t * t);
    }
    s1 = block_sum<double>(s1, red);
    s2 = block_sum<double>(s2, red);
    if (threadIdx.x == 0) { atomicAdd(out, s1); atomicAdd(out + 1, s2); }
}
// grad wrt b:  gb = -(c1*g1*sign(t) + c2*g2*2t) * m
__global__ void masked_diff_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                       const float* __restrict__ m1, const float* __restrict__ m2, int B, int C,
                                       long plane, const float* __restrict__ g1, const float* __restrict__ g2,
                                       float c1, float c2, float* __restrict__ gb) {
    long total = (long)B * C * plane;
    float w1 = c1 * (g1 ? *g1 : 0.f), w2 = c2 * (g2 ? *g2 : 0.f);
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long bc = idx / plane;
        long mo = (bc / C) * plane + (idx - bc * plane);
        float m = m1[mo];
        float ta = a[idx] * m, tb = b[idx] * m;
        if (m2) { float mm = m2[mo]; ta *= mm; tb *= mm; m *= mm; }
        float t = ta - tb;
        float sg = (t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f);
        gb[idx] = -(w1 * sg + w2 * 2.f * t) * m;
    }
}

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
// one pyramid level: d (B,1,h,w), img (B,C,h,w).  'x' = difference along H, 'y' = along W.
// out[0] += sum |dx * wx| ; out[1] += sum |dy * wy|
__device__ __forceinline__ float smooth_weight(const float* img, int C, long plane, long o, long step) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += fabsf(img[c * plane + o] - img[c * plane + o + step]);
    return __expf(-s / (float)C);
}
__global__ void smooth_level_fwd_kernel(const float* __restrict__ d, const float* __restrict__ img, int B, int C,
                                        int h, int w, double* __restrict__ out) {
    __shared__ double red[32];
    long plane = (long)h * w, total = (long)B * plane;
    double sx = 0, sy = 0;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int b = (int)(idx / plane);
        long p = idx - (long)b * plane;
        int i = (int)(p / w), j = (int)(p % w);
        const float* im = img + (long)b * C * plane;
        float v = d[idx];
        if (i < h - 1) sx += (double)fabsf((v - d[idx + w]) * smooth_weight(im, C, plane, p, w));
        if (j < w - 1) sy += (double)fabsf((v - d[idx + 1]) * smooth_weight(im, C, plane, p, 1));
    }
    sx = block_sum<double>(sx, red); sy = block_sum<double>(sy, red);
    if (threadIdx.x == 0) { atomicAdd(out, sx); atomicAdd(out + 1, sy); }
}
// gd[idx] (+)= g * (cx * d/dd sum|dx wx| + cy * d/dd sum|dy wy|)
__global__ void smooth_level_bwd_kernel(const float* __restrict__ d, const float* __restrict__ img, int B, int C,
                                        int h, int w, const float* __restrict__ gscale, float cx, float cy,
                                        float* __restrict__ gd, int accumulate) {
    long plane = (long)h * w, total = (long)B * plane;
    float g = gscale ? *gscale : 1.f;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int b = (int)(idx / plane);
        long p = idx - (long)b * plane;
        int i = (int)(p / w), j = (int)(p % w);
        const float* im = img + (long)b * C * plane;
        float v = d[idx], acc = 0.f;
        if (i < h - 1) { float t = (v - d[idx + w]); float wt = smooth_weight(im, C, plane, p, w);
                         float s = t * wt; acc += cx * wt * ((s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f)); }
        if (i > 0)     { float t = (d[idx - w] - v); float wt = smooth_weight(im, C, plane, p - w, w);
                         float s = t * wt; acc -= cx * wt * ((s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f)); }
        if (j < w - 1) { float t = (v - d[idx + 1]); float wt = smooth_weight(im, C, plane, p, 1);
                         float s = t * wt; acc += cy * wt * ((s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f)); }
        if (j > 0)     { float t = (d[idx - 1] - v); float wt = smooth_weight(im, C, plane, p - 1, 1);
                         float s = t * wt; acc -= cy * wt * ((s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f)); }
        if (accumulate) gd[idx] += g * acc; else gd[idx] = g * acc;
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

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
#define ST(s) ((cudaStream_t)(s))

extern "C" int dsr_hole_valid_masks(const float* depth, int B, int H, int W, float border, float* hole,
                                    float* valid, void* stream) {
    DSR_REQUIRE(depth && valid && B > 0 && H > 0 && W > 0, "bad arguments");
    long n = (long)B * H * W;
    hole_valid_kernel<<<dsr_grid(n, TPB), TPB, 0, ST(stream)>>>(depth, B, H, W, border, hole, valid);
    return dsr_check_launch("hole_valid_masks");
}

extern "C" int dsr_rect_holes(const float* valid, const float* depth, const int* rects, const int* counts,
                              int max_rects, int B, int H, int W, float extra_border, unsigned char* gt_mask,
                              float* masked, float* extra, void* stream) {
    DSR_REQUIRE(valid && depth && rects && counts && gt_mask && masked, "null pointer");
    DSR_REQUIRE(max_rects > 0 && max_rects <= 1024, "max_rects out of range");
    int gx = dsr_cdiv((long)H * W, TPB);
    int cap = dsr_num_sms() * 8 / (B > 0 ? B : 1);
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    dim3 grid(gx, B);
    rect_holes_kernel<<<grid, TPB, max_rects * 4 * sizeof(int), ST(stream)>>>(valid, depth, rects, counts, max_rects,
                                                                              H, W, extra_border, gt_mask, masked, extra);
    return dsr_check_launch("rect_holes");
}

extern "C" int dsr_normals_old_fwd(const float* depth, int B, int H, int W, float scale, float* out, void* stream) {
    DSR_REQUIRE(depth && out && H >= 2 && W >= 2, "bad arguments");
    normals_old_fwd_kernel<<<dsr_grid((long)B * H * W, TPB), TPB, 0, ST(stream)>>>(depth, B, H, W, scale, out);
    return dsr_check_launch("normals_old_fwd");
}
extern "C" int dsr_normals_old_bwd(const float* depth, const float* gout, int B, int H, int W, float scale,
                                   float* gdepth, void* stream) {
    DSR_REQUIRE(depth && gout && gdepth && H >= 2 && W >= 2, "bad arguments");
    normals_old_bwd_kernel<<<dsr_grid((long)B * H * W, TPB), TPB, 0, ST(stream)>>>(depth, gout, B, H, W, scale, gdepth);
    return dsr_check_launch("normals_old_bwd");
}
extern "C" int dsr_normals_new_fwd(const float* depth, const double* cams, int B, int H, int W, float* out,
                                   void* stream) {
    DSR_REQUIRE(depth && cams && out && H >= 2 && W >= 2, "bad arguments");
    normals_new_fwd_kernel<<<dsr_grid((long)B * H * W, TPB), TPB, 0, ST(stream)>>>(depth, cams, B, H, W, out);
    return dsr_check_launch("normals_new_fwd");
}
extern "C" int dsr_normals_new_bwd(const float* depth, const float* gout, const double* cams, int B, int H, int W,
                                   float* gdepth, void* stream) {
    DSR_REQUIRE(depth && gout && cams && gdepth && H >= 2 && W >= 2, "bad arguments");
    normals_new_bwd_kernel<<<dsr_grid((long)B * H * W, TPB), TPB, 0, ST(stream)>>>(depth, gout, cams, B, H, W, gdepth);
    return dsr_check_launch("normals_new_bwd");
}

extern "C" int dsr_tv_fwd(const float* x, long planes, int H, int W, double* out_sum, void* stream) {
    DSR_REQUIRE(x && out_sum, "null pointer");
    tv_fwd_kernel<<<dsr_grid(planes * H * W, TPB), TPB, 0, ST(stream)>>>(x, planes, H, W, out_sum);
    return dsr_check_launch("tv_fwd");
}
extern "C" int dsr_tv_bwd(const float* x, long planes, int H, int W, const float* gscale, float coef, float* gx,
                          void* stream) {
    DSR_REQUIRE(x && gx, "null pointer");
    tv_bwd_kernel<<<dsr_grid(planes * H * W, TPB), TPB, 0, ST(stream)>>>(x, planes, H, W, gscale, coef, gx);
    return dsr_check_launch("tv_bwd");
}

extern "C" int dsr_masked_diff_fwd(const float* a, const float* b, const float* m1, const float* m2, int B, int C,
                                   long plane, double* out2, void* stream) {
    DSR_REQUIRE(a && b && m1 && out2, "null pointer");
    masked_diff_fwd_kernel<<<dsr_grid((long)B * C * plane, TPB), TPB, 0, ST(stream)>>>(a, b, m1, m2, B, C, plane, out2);
    return dsr_check_launch("masked_diff_fwd");
}
extern "C" int dsr_masked_diff_bwd(const float* a, const float* b, const float* m1, const float* m2, int B, int C,
                                   long plane, const float* g_l1, const float* g_l2, float c1, float c2, float* gb,
                                   void* stream) {
    DSR_REQUIRE(a && b && m1 && gb, "null pointer");
    masked_diff_bwd_kernel<<<dsr_grid((long)B * C * plane, TPB), TPB, 0, ST(stream)>>>(a, b, m1, m2, B, C, plane, g_l1,
                                                                                        g_l2, c1, c2, gb);
    return dsr_check_launch("masked_diff_bwd");
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
extern "C" int dsr_smooth_level_fwd(const float* d, const float* img, int B, int C, int h, int w, double* out2,
                                    void* stream) {
    DSR_REQUIRE(d && img && out2, "null pointer");
    smooth_level_fwd_kernel<<<dsr_grid((long)B * h * w, TPB), TPB, 0, ST(stream)>>>(d, img, B, C, h, w, out2);
    return dsr_check_launch("smooth_level_fwd");
}
extern "C" int dsr_smooth_level_bwd(const float* d, const float* img, int B, int C, int h, int w, const float* gscale,
                                    float cx, float cy, float* gd, int accumulate, void* stream) {
    DSR_REQUIRE(d && img && gd, "null pointer");
    smooth_level_bwd_kernel<<<dsr_grid((long)B * h * w, TPB), TPB, 0, ST(stream)>>>(d, img, B, C, h, w, gscale, cx, cy,
                                                                                     gd, accumulate);
    return dsr_check_launch("smooth_level_bwd");
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
    dim3 grid(dsr_cdiv(W, SSIM_T), dsr_cdiv(H, SSIM_T), (unsigned)planes);
    ssim_kernel<<<grid, 256, 0, ST(stream)>>>(a, b, H, W, out_sum, map);
    return dsr_check_launch("ssim_fwd");
}
