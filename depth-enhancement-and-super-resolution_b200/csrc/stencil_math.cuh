// Per-pixel math of the depth-derived stencils, shared by the CUDA kernels (device) and by the
// host-side formula check in tests/ (compiled as plain C++ there).  No memory-layout policy in here:
// every function works on one NCHW plane (H x W, row-major) addressed by (i, j).
//
// Reference behaviour being reproduced (citations into /root/reference):
//   dsr_grad / dsr_gcoef      models/norms.py:115-158   (np.gradient-style differences)
//   old_normal_*              models/norms.py:185-190   (image-space normals, fp32)
//   new_normal_*              models/norms.py:75-108, :29-73 (camera-space normals, fp64 inside)
#pragma once
#include <math.h>

#ifndef DSR_HD
#ifdef __CUDACC__
#define DSR_HD __host__ __device__ __forceinline__
#else
#define DSR_HD inline
#endif
#endif

// derivative of a line f[0..n) (element stride `st`) at index i: central/2 inside, one-sided at ends
template <typename T>
DSR_HD T dsr_grad(const T* f, int i, int n, long st) {
    if (i == 0) return f[st] - f[0];
    if (i == n - 1) return f[(long)(n - 1) * st] - f[(long)(n - 2) * st];
    return (f[(long)(i + 1) * st] - f[(long)(i - 1) * st]) / (T)2;
}
// d grad[j] / d f[i]
DSR_HD float dsr_gcoef(int j, int i, int n) {
    if (j == 0) return i == 1 ? 1.f : (i == 0 ? -1.f : 0.f);
    if (j == n - 1) return i == n - 1 ? 1.f : (i == n - 2 ? -1.f : 0.f);
    return i == j + 1 ? 0.5f : (i == j - 1 ? -0.5f : 0.f);
}

// ---------------------------------------------------------------------------------------------
// old (image-space) normals: n = (-dH, -dW, 1) / (|.| + 1e-6), times `scale`
// ---------------------------------------------------------------------------------------------
DSR_HD void old_normal_fwd(const float* d, int H, int W, int i, int j, float scale, float out[3]) {
    float gh = dsr_grad<float>(d + j, i, H, W);
    float gw = dsr_grad<float>(d + (long)i * W, j, W, 1);
    float v0 = -gh, v1 = -gw, v2 = 1.f;
    float r = sqrtf(v0 * v0 + v1 * v1 + v2 * v2);
    float den = r + 1e-6f;
    out[0] = (v0 / den) * scale;
    out[1] = (v1 / den) * scale;
    out[2] = (v2 / den) * scale;
}
// adjoint at pixel q=(i,j): from dL/dout (3 values at q) to (dL/dgh, dL/dgw) at q
DSR_HD void old_normal_adj(const float* d, int H, int W, int i, int j, float scale,
                           float g0, float g1, float g2, float& dgh, float& dgw) {
    float gh = dsr_grad<float>(d + j, i, H, W);
    float gw = dsr_grad<float>(d + (long)i * W, j, W, 1);
    float v0 = -gh, v1 = -gw, v2 = 1.f;
    float r = sqrtf(v0 * v0 + v1 * v1 + v2 * v2);
    float den = r + 1e-6f;
    float dn0 = g0 * scale, dn1 = g1 * scale, dn2 = g2 * scale;
    float dot = dn0 * v0 + dn1 * v1 + dn2 * v2;
    float k = dot / (r * den * den);
    float dv0 = dn0 / den - k * v0;
    float dv1 = dn1 / den - k * v1;
    dgh = -dv0;
    dgw = -dv1;
}
// dL/dd at (i,j) gathered from the adjoints of the (up to) 5 pixels whose normal reads d(i,j)
DSR_HD float old_normal_bwd(const float* d, const float* g /*3 planes*/, long plane, int H, int W,
                            int i, int j, float scale) {
    float acc = 0.f;
    for (int q = i - 1; q <= i + 1; ++q) {
        if (q < 0 || q >= H) continue;
        float c = dsr_gcoef(q, i, H);
        if (c == 0.f) continue;
        long o = (long)q * W + j;
        float a, b;
        old_normal_adj(d, H, W, q, j, scale, g[o], g[plane + o], g[2 * plane + o], a, b);
        acc += c * a;
    }
    for (int q = j - 1; q <= j + 1; ++q) {
        if (q < 0 || q >= W) continue;
        float c = dsr_gcoef(q, j, W);
        if (c == 0.f) continue;
        long o = (long)i * W + q;
        float a, b;
        old_normal_adj(d, H, W, i, q, scale, g[o], g[plane + o], g[2 * plane + o], a, b);
        acc += c * b;
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// new (camera-space) normals.  cam[0..8] = K^-1 (row major), cam[9] = w0 + shift, cam[10] = h0 + shift
// ---------------------------------------------------------------------------------------------
#define DSR_CAM_DOUBLES 11
DSR_HD void cam_ray(const double* cam, int i, int j, double& rx, double& ry) {
    double u = cam[9] + (double)j, v = cam[10] + (double)i;
    double a = cam[0] * u + cam[1] * v + cam[2];
    double b = cam[3] * u + cam[4] * v + cam[5];
    double c = cam[6] * u + cam[7] * v + cam[8];
    rx = a / c;
    ry = b / c;
}
DSR_HD void cam_point(const float* d, const double* cam, int W, int i, int j, double P[3]) {
    double z = ((double)d[(long)i * W + j] + 1.0) / 2.0;
    double rx, ry;
    cam_ray(cam, i, j, rx, ry);
    P[0] = rx * z;
    P[1] = ry * z;
    P[2] = z;
}
// derivatives of the point map along u (W) and v (H) at (i,j)
DSR_HD void cam_point_grads(const float* d, const double* cam, int H, int W, int i, int j,
                            double Pu[3], double Pv[3]) {
    double A[3], B[3];
    int ja = (j == 0) ? 0 : j - 1, jb = (j == W - 1) ? W - 1 : j + 1;
    if (j == 0) jb = 1;
    if (j == W - 1) ja = W - 2;
    double su = (j == 0 || j == W - 1) ? 1.0 : 2.0;
    cam_point(d, cam, W, i, ja, A);
    cam_point(d, cam, W, i, jb, B);
    for (int c = 0; c < 3; ++c) Pu[c] = (B[c] - A[c]) / su;
    int ia = (i == 0) ? 0 : i - 1, ib = (i == H - 1) ? H - 1 : i + 1;
    if (i == 0) ib = 1;
    if (i == H - 1) ia = H - 2;
    double sv = (i == 0 || i == H - 1) ? 1.0 : 2.0;
    cam_point(d, cam, W, ia, j, A);
    cam_point(d, cam, W, ib, j, B);
    for (int c = 0; c < 3; ++c) Pv[c] = (B[c] - A[c]) / sv;
}
DSR_HD void new_normal_fwd(const float* d, const double* cam, int H, int W, int i, int j, float out[3]) {
    double Pu[3], Pv[3];
    cam_point_grads(d, cam, H, W, i, j, Pu, Pv);
    double m0 = Pv[1] * Pu[2] - Pu[1] * Pv[2];
    double m1 = Pv[2] * Pu[0] - Pu[2] * Pv[0];
    double m2 = Pv[0] * Pu[1] - Pu[0] * Pv[1];
    double r = sqrt(m0 * m0 + m1 * m1 + m2 * m2);
    double den = r > 1e-12 ? r : 1e-12;
    out[0] = (float)(m0 / den);
    out[1] = (float)(m1 / den);
    out[2] = (float)(m2 / den);
}
// adjoint at q=(i,j): dL/dn (3) -> dL/dPu, dL/dPv at q
DSR_HD void new_normal_adj(const float* d, const double* cam, int H, int W, int i, int j,
                           double g0, double g1, double g2, double dPu[3], double dPv[3]) {
    double Pu[3], Pv[3];
    cam_point_grads(d, cam, H, W, i, j, Pu, Pv);
    double m[3] = {Pv[1] * Pu[2] - Pu[1] * Pv[2], Pv[2] * Pu[0] - Pu[2] * Pv[0], Pv[0] * Pu[1] - Pu[0] * Pv[1]};
    double r = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    double dm[3];
    if (r > 1e-12) {
        double n0 = m[0] / r, n1 = m[1] / r, n2 = m[2] / r;
        double dot = n0 * g0 + n1 * g1 + n2 * g2;
        dm[0] = (g0 - n0 * dot) / r;
        dm[1] = (g1 - n1 * dot) / r;
        dm[2] = (g2 - n2 * dot) / r;
    } else {  // clamped denominator: n = m / 1e-12
        dm[0] = g0 / 1e-12; dm[1] = g1 / 1e-12; dm[2] = g2 / 1e-12;
    }
    // m = Pv x Pu  =>  dPv = Pu x dm,  dPu = dm x Pv
    dPv[0] = Pu[1] * dm[2] - Pu[2] * dm[1];
    dPv[1] = Pu[2] * dm[0] - Pu[0] * dm[2];
    dPv[2] = Pu[0] * dm[1] - Pu[1] * dm[0];
    dPu[0] = dm[1] * Pv[2] - dm[2] * Pv[1];
    dPu[1] = dm[2] * Pv[0] - dm[0] * Pv[2];
    dPu[2] = dm[0] * Pv[1] - dm[1] * Pv[0];
}
DSR_HD float new_normal_bwd(const float* d, const float* g /*3 planes*/, long plane, const double* cam,
                            int H, int W, int i, int j) {
    double rx, ry;
    cam_ray(cam, i, j, rx, ry);
    double acc = 0.0;  // dL/dz at (i,j)
    for (int q = j - 1; q <= j + 1; ++q) {
        if (q < 0 || q >= W) continue;
        double c = (double)dsr_gcoef(q, j, W);
        if (c == 0.0) continue;
        long o = (long)i * W + q;
        double dPu[3], dPv[3];
        new_normal_adj(d, cam, H, W, i, q, (double)g[o], (double)g[plane + o], (double)g[2 * plane + o], dPu, dPv);
        acc += c * (dPu[0] * rx + dPu[1] * ry + dPu[2]);
    }
    for (int q = i - 1; q <= i + 1; ++q) {
        if (q < 0 || q >= H) continue;
        double c = (double)dsr_gcoef(q, i, H);
        if (c == 0.0) continue;
        long o = (long)q * W + j;
        double dPu[3], dPv[3];
        new_normal_adj(d, cam, H, W, q, j, (double)g[o], (double)g[plane + o], (double)g[2 * plane + o], dPu, dPv);
        acc += c * (dPv[0] * rx + dPv[1] * ry + dPv[2]);
    }
    return (float)(acc * 0.5);  // z = (d + 1) / 2
}

// --- adjoint of the camera-space normal in fp32, from fp32 copies of the fp64 point differences (the tiled kernels of
// stencil_tiled.cu: differences and the forward cross product stay fp64 - both cancel - the rest is fp32)
DSR_HD void new_normal_adj_from_grads(const float Pu[3], const float Pv[3], float g0, float g1, float g2,
                                      float dPu[3], float dPv[3]) {
    float m[3] = {Pv[1] * Pu[2] - Pu[1] * Pv[2], Pv[2] * Pu[0] - Pu[2] * Pv[0], Pv[0] * Pu[1] - Pu[0] * Pv[1]};
    float r = sqrtf(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    float dm[3];
    if (r > 1e-12f) {
        float ir = 1.f / r;
        float n0 = m[0] * ir, n1 = m[1] * ir, n2 = m[2] * ir;
        float dot = n0 * g0 + n1 * g1 + n2 * g2;
        dm[0] = (g0 - n0 * dot) * ir; dm[1] = (g1 - n1 * dot) * ir; dm[2] = (g2 - n2 * dot) * ir;
    } else {
        dm[0] = g0 / 1e-12f; dm[1] = g1 / 1e-12f; dm[2] = g2 / 1e-12f;
    }
    dPv[0] = Pu[1] * dm[2] - Pu[2] * dm[1];
    dPv[1] = Pu[2] * dm[0] - Pu[0] * dm[2];
    dPv[2] = Pu[0] * dm[1] - Pu[1] * dm[0];
    dPu[0] = dm[1] * Pv[2] - dm[2] * Pv[1];
    dPu[1] = dm[2] * Pv[0] - dm[0] * Pv[2];
    dPu[2] = dm[0] * Pv[1] - dm[1] * Pv[0];
}

// ---------------------------------------------------------------------------------------------
// Camera-space normals for AFFINE rays (last row of K^-1 == (0, 0, 1): every pin-hole K), closed form in fp32.
// With ray(i, j) = (rx, ry, 1), rx = k0 u + k1 v + k2, ry = k3 u + k4 v + k5, the neighbours' rays are rx +- k0 / k1, so
//   P(j+1) - P(j-1) = (rx Du + k0 Su, ry Du + k3 Su, Du),  Du = z_r - z_l,  Su = mb z_r + ma z_l   (ma / mb = 0 on the border
//   P(i+1) - P(i-1) = (rx Dv + k1 Sv, ry Dv + k4 Sv, Dv),  Dv = z_d - z_u,  Sv = md z_d + mu z_u    whose neighbour is clamped)
// and in m = Pv x Pu the rx ry Du Dv products cancel ANALYTICALLY:
//   m0 = k4 Sv Du - k3 Su Dv,   m1 = k0 Su Dv - k1 Sv Du,   m2 = -rx m0 - ry m1 + (k1 k3 - k0 k4) Su Sv
// Du / Dv are differences of fp32 depths (exact to an ulp), so nothing cancels numerically any more and the fp64 detour of
// the reference (norms.py:85-108) is not needed: ~35 fp32 instructions per pixel instead of ~50 fp64 ones.  `sc` = the
// product of np.gradient's 1/2 (interior) or 1 (border) factors; z = (d + 1) / 2.
// ---------------------------------------------------------------------------------------------
struct AffCam {
    float k0, k1, k3, k4, D;
};
DSR_HD bool cam_is_affine(const double* cam) { return cam[6] == 0.0 && cam[7] == 0.0 && cam[8] == 1.0; }
DSR_HD AffCam aff_cam(const double* cam) {
    AffCam a;
    a.k0 = (float)cam[0]; a.k1 = (float)cam[1]; a.k3 = (float)cam[3]; a.k4 = (float)cam[4];
    a.D = (float)(cam[1] * cam[3] - cam[0] * cam[4]);
    return a;
}
// the four bilinear building blocks of one pixel from its (border-clamped) depth neighbours
DSR_HD void aff_terms(float dl, float dr, float du, float dd, float ma, float mb, float mu, float md,
                      float& Du, float& Su, float& Dv, float& Sv) {
    Du = 0.5f * (dr - dl);
    Su = 0.5f * (mb * (dr + 1.f) + ma * (dl + 1.f));
    Dv = 0.5f * (dd - du);
    Sv = 0.5f * (md * (dd + 1.f) + mu * (du + 1.f));
}
DSR_HD void aff_normal_m(const AffCam& c, float rx, float ry, float Du, float Su, float Dv, float Sv, float sc, float m[3]) {
    const float A = Su * Dv, B = Sv * Du;
    m[0] = sc * (c.k4 * B - c.k3 * A);
    m[1] = sc * (c.k0 * A - c.k1 * B);
    m[2] = sc * (c.D * Su * Sv) - rx * m[0] - ry * m[1];
}
// 1 / max(|m|, 1e-12) = rsqrt(max(|m|^2, 1e-24)): one MUFU on the device (2 ulp; the parity gate on normals is 1e-6 absolute)
DSR_HD float aff_inv_len(float r2) {
#ifdef __CUDA_ARCH__
    return rsqrtf(fmaxf(r2, 1e-24f));
#else
    const float r = sqrtf(r2);
    return 1.f / (r > 1e-12f ? r : 1e-12f);
#endif
}
DSR_HD void aff_normalize(const float m[3], float n[3]) {
    const float k = aff_inv_len(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    n[0] = m[0] * k; n[1] = m[1] * k; n[2] = m[2] * k;
}
// adjoint of one pixel: dL/dn (g) -> what it adds to dL/dd of its right / left / lower / upper neighbour (R, L, Dn, Up)
DSR_HD void aff_pixel_adj(const AffCam& c, float rx, float ry, float dl, float dr, float du, float dd, float ma, float mb,
                          float mu, float md, float sc, float g0, float g1, float g2, float& R, float& L, float& Dn, float& Up) {
    float Du, Su, Dv, Sv, m[3], dm[3];
    aff_terms(dl, dr, du, dd, ma, mb, mu, md, Du, Su, Dv, Sv);
    aff_normal_m(c, rx, ry, Du, Su, Dv, Sv, sc, m);
    const float r2 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
    const float ir = aff_inv_len(r2);
    if (r2 > 1e-24f) {
        const float n0 = m[0] * ir, n1 = m[1] * ir, n2 = m[2] * ir;
        const float dot = n0 * g0 + n1 * g1 + n2 * g2;
        dm[0] = (g0 - n0 * dot) * ir; dm[1] = (g1 - n1 * dot) * ir; dm[2] = (g2 - n2 * dot) * ir;
    } else {                                   // clamped denominator: n = m / 1e-12
        dm[0] = g0 * 1e12f; dm[1] = g1 * 1e12f; dm[2] = g2 * 1e12f;
    }
    const float e0 = dm[0] - rx * dm[2], e1 = dm[1] - ry * dm[2], e2 = sc * c.D * dm[2];
    const float p = sc * (c.k4 * e0 - c.k1 * e1), q = sc * (c.k0 * e1 - c.k3 * e0);
    const float dDu = Sv * p, dSv = Du * p + e2 * Su, dDv = Su * q, dSu = Dv * q + e2 * Sv;
    R = 0.5f * (dDu + mb * dSu);
    L = 0.5f * (ma * dSu - dDu);
    Dn = 0.5f * (dDv + md * dSv);
    Up = 0.5f * (mu * dSv - dDv);
}
// whole-plane helpers (host check / reference form of what the tiled kernels compute)
DSR_HD void aff_pixel_setup(const float* d, const double* cam, int H, int W, int i, int j, float& rx, float& ry, float& dl, float& dr,
                            float& du, float& dd, float& ma, float& mb, float& mu, float& md, float& sc) {
    double rxd, ryd;
    cam_ray(cam, i, j, rxd, ryd);
    rx = (float)rxd; ry = (float)ryd;
    const float* row = d + (long)i * W;
    dl = row[j > 0 ? j - 1 : 0]; dr = row[j < W - 1 ? j + 1 : W - 1];
    du = d[(long)(i > 0 ? i - 1 : 0) * W + j]; dd = d[(long)(i < H - 1 ? i + 1 : H - 1) * W + j];
    ma = j > 0 ? 1.f : 0.f; mb = j < W - 1 ? 1.f : 0.f; mu = i > 0 ? 1.f : 0.f; md = i < H - 1 ? 1.f : 0.f;
    sc = ((j == 0 || j == W - 1) ? 1.f : 0.5f) * ((i == 0 || i == H - 1) ? 1.f : 0.5f);
}
DSR_HD void aff_normal_fwd(const float* d, const double* cam, int H, int W, int i, int j, float out[3]) {
    float rx, ry, dl, dr, du, dd, ma, mb, mu, md, sc, Du, Su, Dv, Sv, m[3];
    aff_pixel_setup(d, cam, H, W, i, j, rx, ry, dl, dr, du, dd, ma, mb, mu, md, sc);
    aff_terms(dl, dr, du, dd, ma, mb, mu, md, Du, Su, Dv, Sv);
    aff_normal_m(aff_cam(cam), rx, ry, Du, Su, Dv, Sv, sc, m);
    aff_normalize(m, out);
}
DSR_HD float aff_normal_bwd(const float* d, const float* g /*3 planes*/, long plane, const double* cam, int H, int W, int i, int j) {
    const AffCam c = aff_cam(cam);
    float acc = 0.f;
    // pixels whose right / left / lower / upper (border-clamped) neighbour is (i, j)
    for (int t = 0; t < 5; ++t) {
        const int qi = i + (t == 3 ? -1 : (t == 4 ? 1 : 0)), qj = j + (t == 1 ? -1 : (t == 2 ? 1 : 0));
        if (qi < 0 || qi >= H || qj < 0 || qj >= W) continue;
        float rx, ry, dl, dr, du, dd, ma, mb, mu, md, sc, R, L, Dn, Up;
        aff_pixel_setup(d, cam, H, W, qi, qj, rx, ry, dl, dr, du, dd, ma, mb, mu, md, sc);
        const long o = (long)qi * W + qj;
        aff_pixel_adj(c, rx, ry, dl, dr, du, dd, ma, mb, mu, md, sc, g[o], g[plane + o], g[2 * plane + o], R, L, Dn, Up);
        if (t == 0) acc += (j == W - 1 ? R : 0.f) + (j == 0 ? L : 0.f) + (i == H - 1 ? Dn : 0.f) + (i == 0 ? Up : 0.f);
        else if (t == 1) acc += R;
        else if (t == 2) acc += L;
        else if (t == 3) acc += Dn;
        else acc += Up;
    }
    return acc;
}

// bilinear, align_corners=True (torch upsample_bilinear2d): source index and lambda for output o
DSR_HD void bilin_ac(int o, int n_out, int n_in, int& i0, int& i1, float& l0, float& l1) {
    float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
    float src = scale * (float)o;
    i0 = (int)src;
    if (i0 > n_in - 1) i0 = n_in - 1;
    i1 = i0 + ((i0 < n_in - 1) ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}
