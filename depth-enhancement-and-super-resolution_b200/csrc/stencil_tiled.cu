// Tiled forms of the depth-derived stencils of the loss stack (HBM-bound; NCHW fp32 planes as the reference holds them).
// Reference call sites: models/main_model.py:208-230 (masks), :257-298 (rectangle holes), :15-19 (TV), :22-73
// (smoothness), :340-417 (masked L1 / MSE terms), models/norms.py:75-235 (normals).
//
// Why tiles: the first-generation kernels spent their time on 64-bit index divisions and (camera-space normals) on
// ~50 fp64 divisions per pixel, not on memory (profiles/r1b_stencils_before.json: 2-35 % of HBM peak).  Here one CTA
// of 256 threads owns a 16 x 64 pixel tile of one plane: the tile (+ halo) is staged once in shared memory with
// coalesced loads, every thread then produces 4 horizontally adjacent pixels (one 16-byte store per output plane), all
// index arithmetic is 32-bit with compile-time divisors, and intermediate per-pixel quantities that several output
// pixels share (the fp64 point map, the per-pixel adjoints of the normal) are computed once per pixel in shared memory.
// CTAs walk the tile list with a grid-stride loop (grid = min(tiles, 8 x SMs)).
#include <stdlib.h>
#include "common.cuh"
#include "stencil_math.cuh"
#include "../../include/dsr_b200.h"

#define ST(s) ((cudaStream_t)(s))
#define TH 16
#define TW 64
#define NT 256

struct TileIter {                        // decomposition of a linear tile index (32-bit divisions, once per tile)
    int tiles_w, tiles_h, ntiles;
    __host__ __device__ TileIter(int planes, int H, int W)
        : tiles_w((W + TW - 1) / TW), tiles_h((H + TH - 1) / TH), ntiles(planes * tiles_h * tiles_w) {}
    __device__ void at(int t, int& pl, int& i0, int& j0) const {
        const int tw = t % tiles_w; t /= tiles_w;
        const int th = t % tiles_h; pl = t / tiles_h;
        i0 = th * TH; j0 = tw * TW;
    }
};
static int tile_grid(int planes, int H, int W) {
    TileIter it(planes, H, W);
    const long cap = (long)dsr_num_sms() * 8;
    return (int)(it.ntiles < cap ? it.ntiles : cap);
}

// 4 adjacent outputs of row i starting at column j (j % 4 == 0): one 16-byte store when the row allows it
__device__ __forceinline__ void store4(float* __restrict__ plane, int W, int i, int j, const float v[4], bool vec) {
    float* o = plane + (long)i * W + j;
    if (vec && j + 3 < W) { st4(o, make_float4(v[0], v[1], v[2], v[3])); return; }
#pragma unroll
    for (int e = 0; e < 4; ++e) if (j + e < W) o[e] = v[e];
}
__device__ __forceinline__ bool vec_ok(const void* p, int W) { return (W & 3) == 0 && ((uintptr_t)p & 15) == 0; }

// ------------------------------------------------------------------------------------------
// Register-quad kernels (no shared memory, no divisions): grid = (W tiles, H tiles, planes), thread (ty, tx) owns the
// 4 pixels (i0 + ty, j0 + 4 tx .. + 3).  Neighbour rows / columns come from clamped loads (L1 hits: the rows above and
// below are the centre rows of the neighbouring threads), which also encodes np.gradient's one-sided differences:
// d/di = (x[min(i+1, H-1)] - x[max(i-1, 0)]) * (border ? 1 : 0.5).
// ------------------------------------------------------------------------------------------
struct Quad {
    int pl, i, j;
    bool ok;
    __device__ __forceinline__ Quad(int H, int W) {
        pl = blockIdx.z;
        i = blockIdx.y * TH + (threadIdx.x >> 4);
        j = blockIdx.x * TW + ((threadIdx.x & 15) << 2);
        ok = i < H && j < W;
    }
};
static dim3 quad_grid(int planes, int H, int W) { return dim3((W + TW - 1) / TW, (H + TH - 1) / TH, planes); }
// v[0..5] = x[j-1 .. j+4] of one row, columns clamped into [0, W)
__device__ __forceinline__ void load6(const float* __restrict__ row, int j, int W, bool vec, float v[6]) {
    if (vec) {
        const float4 c = ld4(row + j);
        v[1] = c.x; v[2] = c.y; v[3] = c.z; v[4] = c.w;
        v[0] = j > 0 ? __ldg(row + j - 1) : c.x;
        v[5] = j + 4 < W ? __ldg(row + j + 4) : c.w;
    } else {
#pragma unroll
        for (int e = 0; e < 6; ++e) v[e] = __ldg(row + min(max(j - 1 + e, 0), W - 1));
    }
}
__device__ __forceinline__ void load4(const float* __restrict__ row, int j, int W, bool vec, float v[4]) {
    if (vec) {
        const float4 c = ld4(row + j);
        v[0] = c.x; v[1] = c.y; v[2] = c.z; v[3] = c.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = __ldg(row + min(j + e, W - 1));
    }
}
__device__ __forceinline__ float edge_f(int i, int n) { return (i == 0 || i == n - 1) ? 1.f : 0.5f; }

// ------------------------------------------------------------------------------------------
// hole / valid masks (main_model.py:208-230): hole = d <= border; valid = !dilate3x3(hole)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
hole_valid_quad(const float* __restrict__ d, int H, int W, float border, float* __restrict__ hole, float* __restrict__ valid) {
    const Quad q(H, W);
    if (!q.ok) return;
    const long base = (long)q.pl * H * W;
    const bool vin = vec_ok(d, W), vout = vec_ok(valid, W) && (!hole || vec_ok(hole, W));
    float u[6], c[6], l[6];
    load6(d + base + (long)max(q.i - 1, 0) * W, q.j, W, vin, u);       // clamped rows / columns duplicate in-range
    load6(d + base + (long)q.i * W, q.j, W, vin, c);                   // neighbours: the 3x3 OR is unchanged
    load6(d + base + (long)min(q.i + 1, H - 1) * W, q.j, W, vin, l);
    bool col[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) col[e] = (u[e] <= border) | (c[e] <= border) | (l[e] <= border);
    float hv[4], vv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        hv[e] = (c[e + 1] <= border) ? 1.f : 0.f;
        vv[e] = (col[e] | col[e + 1] | col[e + 2]) ? 0.f : 1.f;
    }
    if (hole) store4(hole + base, W, q.i, q.j, hv, vout);
    store4(valid + base, W, q.i, q.j, vv, vout);
}

// ------------------------------------------------------------------------------------------
// rectangle holes (main_model.py:257-298 + :354-357 / :396).  rects int32 [B][max_rects][4] = x, y, sx, sy
//   gt = !(valid > 0.05 && covered);  masked = gt ? depth : -1;  extra = (masked < extra_border) || !gt
// Each CTA first keeps only the rectangles that touch its tile (usually 0-3 of the up-to-59).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
rect_holes_tiled(const float* __restrict__ valid, const float* __restrict__ depth, const int* __restrict__ rects,
                 const int* __restrict__ counts, int max_rects, int B, int H, int W, float extra_border,
                 unsigned char* __restrict__ gt, float* __restrict__ masked, float* __restrict__ extra) {
    extern __shared__ __align__(16) int srect[];   // [max_rects][4] compacted
    __shared__ int s_n;
    const TileIter it(B, H, W);
    const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) << 2;
    const bool vec = vec_ok(masked, W) && vec_ok(valid, W) && vec_ok(depth, W) && (!extra || vec_ok(extra, W)) &&
                     ((uintptr_t)gt & 3) == 0;
    for (int t = blockIdx.x; t < it.ntiles; t += gridDim.x) {
        int b, i0, j0;
        it.at(t, b, i0, j0);
        __syncthreads();
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const int n = counts[b];
        for (int r = threadIdx.x; r < n; r += NT) {
            const int4 q = *reinterpret_cast<const int4*>(rects + ((long)b * max_rects + r) * 4);
            if (q.z > 0 && q.w > 0 && q.x < j0 + TW && q.x + q.z > j0 && q.y < i0 + TH && q.y + q.w > i0) {
                const int k = atomicAdd(&s_n, 1);
                *reinterpret_cast<int4*>(srect + 4 * k) = q;
            }
        }
        __syncthreads();
        const int i = i0 + ty, j = j0 + tx;
        if (i >= H || j >= W) continue;
        const int m = s_n;
        const long o = ((long)b * H + i) * W + j;
        float vv[4], dd[4];
        if (vec && j + 3 < W) {
            const float4 a = ld4(valid + o), c = ld4(depth + o);
            vv[0] = a.x; vv[1] = a.y; vv[2] = a.z; vv[3] = a.w; dd[0] = c.x; dd[1] = c.y; dd[2] = c.z; dd[3] = c.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) { const bool in = j + e < W; vv[e] = in ? valid[o + e] : 0.f; dd[e] = in ? depth[o + e] : 0.f; }
        }
        bool cov[4] = {false, false, false, false};
        for (int r = 0; r < m; ++r) {
            const int rx = srect[4 * r], ry = srect[4 * r + 1], sx = srect[4 * r + 2], sy = srect[4 * r + 3];
            const bool rowin = (i >= ry) & (i < ry + sy);
#pragma unroll
            for (int e = 0; e < 4; ++e) cov[e] |= rowin & (j + e >= rx) & (j + e < rx + sx);
        }
        float mv[4], ev[4];
        unsigned gpack = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool g = !((vv[e] > 0.05f) && cov[e]);
            mv[e] = g ? dd[e] : -1.f;
            ev[e] = ((mv[e] < extra_border) || !g) ? 1.f : 0.f;
            gpack |= (g ? 1u : 0u) << (8 * e);
        }
        if (vec && j + 3 < W) *reinterpret_cast<unsigned*>(gt + o) = gpack;
        else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (j + e < W) gt[o + e] = (unsigned char)((gpack >> (8 * e)) & 1u);
        }
        store4(masked + (long)b * H * W, W, i, j, mv, vec);
        if (extra) store4(extra + (long)b * H * W, W, i, j, ev, vec);
    }
}

// ------------------------------------------------------------------------------------------
// image-space normals (norms.py:185-235)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
normals_old_fwd_quad(const float* __restrict__ d, int H, int W, float scale, float* __restrict__ out) {
    const Quad q(H, W);
    if (!q.ok) return;
    const long plane = (long)H * W;
    const float* p = d + q.pl * plane;
    const bool vin = vec_ok(d, W), vout = vec_ok(out, W);
    float u[4], l[4], c[6];
    load4(p + (long)max(q.i - 1, 0) * W, q.j, W, vin, u);
    load4(p + (long)min(q.i + 1, H - 1) * W, q.j, W, vin, l);
    load6(p + (long)q.i * W, q.j, W, vin, c);
    const float fh = edge_f(q.i, H);
    float n0[4], n1[4], n2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float gh = (l[e] - u[e]) * fh, gw = (c[e + 2] - c[e]) * edge_f(q.j + e, W);
        const float r = sqrtf(gh * gh + gw * gw + 1.f);
        const float k = scale / (r + 1e-6f);            // (v / den) * scale of norms.py:190 up to one rounding
        n0[e] = -gh * k; n1[e] = -gw * k; n2[e] = k;
    }
    float* o = out + (long)q.pl * 3 * plane;
    store4(o, W, q.i, q.j, n0, vout); store4(o + plane, W, q.i, q.j, n1, vout); store4(o + 2 * plane, W, q.i, q.j, n2, vout);
}
// adjoint of one pixel's normal: (dL/dgh, dL/dgw) from dL/dn (fast reciprocal / rsqrt: ~2 ulp, the gradient gate is 1e-4)
__device__ __forceinline__ void old_adj_fast(float gh, float gw, float scale, float g0, float g1, float g2, float& dgh, float& dgw) {
    const float r2 = gh * gh + gw * gw + 1.f;
    const float ir = rsqrtf(r2), r = r2 * ir;
    const float iden = __fdividef(1.f, r + 1e-6f);
    const float dn0 = g0 * scale, dn1 = g1 * scale, dn2 = g2 * scale;
    const float dot = dn2 - dn0 * gh - dn1 * gw;                 // dn . v with v = (-gh, -gw, 1)
    const float k = dot * ir * iden * iden;
    dgh = -(dn0 * iden + k * gh);                                // -(dn0/den - k v0), v0 = -gh
    dgw = -(dn1 * iden + k * gw);
}
// coefficients of np.gradient's adjoint at index i of a line of n: out(i) = cu*A(i-1) + cs*A(i) + cd*A(i+1)
// (A = 0 outside the line)
__device__ __forceinline__ void adj_taps(int i, int n, float& cu, float& cs, float& cd) {
    cu = (i == 1) ? 1.f : 0.5f;
    cd = (i == n - 2) ? -1.f : -0.5f;
    cs = (i == 0) ? -1.f : ((i == n - 1) ? 1.f : 0.f);
}
// One CTA = one 16 x 64 tile.  Phase 1: every thread computes the adjoints of its own 4 pixels from vector loads and
// parks them in shared memory; 64 threads also do the halo (the rows above / below the tile, the columns left / right
// of it).  Phase 2: gather through np.gradient's taps.  Shared tile: row r <-> i0 - 1 + r, column c <-> j0 - 4 + c.
#define AW (TW + 8)
__device__ __forceinline__ void old_adj_quad(const float* __restrict__ p, const float* __restrict__ gp, long plane, int H, int W,
                                             int i, int j, float scale, bool vin, float* __restrict__ sh, float* __restrict__ sw) {
    float dh[4] = {0.f, 0.f, 0.f, 0.f}, dw[4] = {0.f, 0.f, 0.f, 0.f};
    if (i >= 0 && i < H && j < W) {
        float u[4], l[4], c[6], g0[4], g1[4], g2[4];
        load4(p + (long)max(i - 1, 0) * W, j, W, vin, u);
        load4(p + (long)min(i + 1, H - 1) * W, j, W, vin, l);
        load6(p + (long)i * W, j, W, vin, c);
        load4(gp + (long)i * W, j, W, vin, g0);
        load4(gp + plane + (long)i * W, j, W, vin, g1);
        load4(gp + 2 * plane + (long)i * W, j, W, vin, g2);
        const float fh = edge_f(i, H);
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (j + e < W) old_adj_fast((l[e] - u[e]) * fh, (c[e + 2] - c[e]) * edge_f(j + e, W), scale, g0[e], g1[e], g2[e], dh[e], dw[e]);
    }
    *reinterpret_cast<float4*>(sh) = make_float4(dh[0], dh[1], dh[2], dh[3]);
    *reinterpret_cast<float4*>(sw) = make_float4(dw[0], dw[1], dw[2], dw[3]);
}
__global__ void __launch_bounds__(NT)
normals_old_bwd_quad(const float* __restrict__ d, const float* __restrict__ g, int H, int W, float scale, float* __restrict__ gd) {
    __shared__ __align__(16) float a_h[(TH + 2) * AW], a_w[(TH + 2) * AW];
    const int i0 = blockIdx.y * TH, j0 = blockIdx.x * TW;
    const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) << 2;
    const long plane = (long)H * W;
    const float* p = d + blockIdx.z * plane;
    const float* gp = g + (long)blockIdx.z * 3 * plane;
    const bool vin = vec_ok(d, W) && vec_ok(g, W), vout = vec_ok(gd, W);
    old_adj_quad(p, gp, plane, H, W, i0 + ty, j0 + tx, scale, vin, a_h + (ty + 1) * AW + 4 + tx, a_w + (ty + 1) * AW + 4 + tx);
    if (threadIdx.x < 32) {                       // rows i0 - 1 and i0 + TH
        const int r = (threadIdx.x >> 4) ? TH + 1 : 0;
        old_adj_quad(p, gp, plane, H, W, i0 - 1 + r, j0 + tx, scale, vin, a_h + r * AW + 4 + tx, a_w + r * AW + 4 + tx);
    } else if (threadIdx.x < 64) {                // columns j0 - 1 and j0 + TW of the tile's rows (only a_w is read there)
        const int t = threadIdx.x - 32, r = 1 + (t & 15), c = (t >> 4) ? TW + 4 : 3;
        const int i = i0 - 1 + r, j = j0 - 4 + c;
        float dh = 0.f, dw = 0.f;
        if (i < H && j >= 0 && j < W) {
            const float gh = (__ldg(p + (long)min(i + 1, H - 1) * W + j) - __ldg(p + (long)max(i - 1, 0) * W + j)) * edge_f(i, H);
            const float gw = (__ldg(p + (long)i * W + min(j + 1, W - 1)) - __ldg(p + (long)i * W + max(j - 1, 0))) * edge_f(j, W);
            const long o = (long)i * W + j;
            old_adj_fast(gh, gw, scale, __ldg(gp + o), __ldg(gp + plane + o), __ldg(gp + 2 * plane + o), dh, dw);
        }
        a_w[r * AW + c] = dw;
    }
    __syncthreads();
    const int i = i0 + ty, j = j0 + tx;
    if (i >= H || j >= W) return;
    float cu, cs, cd;
    adj_taps(i, H, cu, cs, cd);
    const float* rh = a_h + (ty + 1) * AW + 4 + tx;
    const float* rw = a_w + (ty + 1) * AW + 4 + tx;
    const float4 hu = *reinterpret_cast<const float4*>(rh - AW), hc = *reinterpret_cast<const float4*>(rh),
                 hd = *reinterpret_cast<const float4*>(rh + AW), wc = *reinterpret_cast<const float4*>(rw);
    const float hU[4] = {hu.x, hu.y, hu.z, hu.w}, hC[4] = {hc.x, hc.y, hc.z, hc.w}, hD[4] = {hd.x, hd.y, hd.z, hd.w};
    const float wv[6] = {rw[-1], wc.x, wc.y, wc.z, wc.w, rw[4]};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float wl, ws, wr;
        adj_taps(j + e, W, wl, ws, wr);
        o[e] = cu * hU[e] + cs * hC[e] + cd * hD[e] + wl * wv[e] + ws * wv[e + 1] + wr * wv[e + 2];
    }
    store4(gd + blockIdx.z * plane, W, i, j, o, vout);
}

// ------------------------------------------------------------------------------------------
// camera-space normals, AFFINE rays (every pin-hole K: last row of K^-1 == (0, 0, 1)): the closed fp32 form of
// stencil_math.cuh (aff_*) - no fp64 point map, the rx ry Du Dv products of the cross product cancelled analytically.
// The fp64 tiled kernels this replaces ran at 31 % / 16 % of the HBM peak, bound by ~50 fp64 instructions per pixel (forward)
// and ~280 instructions per pixel (backward, ncu r36).  A camera with a perspective row in K^-1 (never a pin-hole K) takes
// the per-pixel fp64 functions of stencil_math.cuh instead - slow, exact, and chosen per plane ON THE DEVICE (the
// `affine` flag of its camera row), so the choice never depends on host knowledge of K and a captured CUDA graph stays
// valid whatever K the next batch brings.  (Launching a second, separate kernel for those planes cost ~40-80 us of empty
// blocks per call at 96 x 512 x 640 - measured r87 - hence one kernel with a block-uniform branch; the register caps keep
// the occupancy of the fast path and let the rare path spill.)
// ------------------------------------------------------------------------------------------
struct AffPlane {
    AffCam c;
    float rx0, ry0;           // ray of pixel (i, j0q): the quad's pixels are rx0 + e * k0, ry0 + e * k3
    bool affine;
    __device__ __forceinline__ AffPlane(const double* __restrict__ cam) {
        affine = cam_is_affine(cam);
        c = aff_cam(cam);
    }
    __device__ __forceinline__ void at(const double* __restrict__ cam, int i, int j) {
        // affine rays: c = k6 u + k7 v + k8 == 1, so cam_ray()'s two fp64 divisions (~100 instructions per quad) are skipped
        const double u = cam[9] + (double)j, v = cam[10] + (double)i;
        rx0 = (float)(cam[0] * u + cam[1] * v + cam[2]);
        ry0 = (float)(cam[3] * u + cam[4] * v + cam[5]);
    }
};
__device__ __noinline__ void normals_generic_fwd_px(const float* __restrict__ p, const double* __restrict__ cam, int H, int W, int i, int j,
                                                    long plane, float* __restrict__ o) {
    float n[3];
    new_normal_fwd(p, cam, H, W, i, j, n);
    o[(long)i * W + j] = n[0]; o[plane + (long)i * W + j] = n[1]; o[2 * plane + (long)i * W + j] = n[2];
}
__global__ void __launch_bounds__(NT, 4)
normals_new_fwd_quad(const float* __restrict__ d, const double* __restrict__ cams, int H, int W, float* __restrict__ out) {
    const Quad q(H, W);
    const double* cam = cams + blockIdx.z * DSR_CAM_DOUBLES;
    AffPlane a(cam);
    if (!q.ok) return;
    const long plane = (long)H * W;
    const float* p = d + q.pl * plane;
    if (!a.affine) {                                          // perspective row in K^-1: per-pixel fp64 reference form
        for (int e = 0; e < 4 && q.j + e < W; ++e) normals_generic_fwd_px(p, cam, H, W, q.i, q.j + e, plane, out + (long)q.pl * 3 * plane);
        return;
    }
    const bool vin = vec_ok(d, W), vout = vec_ok(out, W);
    float u[4], l[4], c[6];
    load4(p + (long)max(q.i - 1, 0) * W, q.j, W, vin, u);
    load4(p + (long)min(q.i + 1, H - 1) * W, q.j, W, vin, l);
    load6(p + (long)q.i * W, q.j, W, vin, c);
    a.at(cam, q.i, q.j);
    float n0[4], n1[4], n2[4];
    if (q.i > 0 && q.i < H - 1 && q.j > 0 && q.j + 4 < W) {
        // quads that touch no image border (all but a frame of tiles): masks and the 1/2 factors are literals
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float Du, Su, Dv, Sv, m[3], n[3];
            aff_terms(c[e], c[e + 2], u[e], l[e], 1.f, 1.f, 1.f, 1.f, Du, Su, Dv, Sv);
            aff_normal_m(a.c, a.rx0 + (float)e * a.c.k0, a.ry0 + (float)e * a.c.k3, Du, Su, Dv, Sv, 0.25f, m);
            aff_normalize(m, n);
            n0[e] = n[0]; n1[e] = n[1]; n2[e] = n[2];
        }
    } else {
        const float mu = q.i > 0 ? 1.f : 0.f, md = q.i < H - 1 ? 1.f : 0.f, fh = edge_f(q.i, H);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = q.j + e;
            float Du, Su, Dv, Sv, m[3], n[3];
            aff_terms(c[e], c[e + 2], u[e], l[e], j > 0 ? 1.f : 0.f, j < W - 1 ? 1.f : 0.f, mu, md, Du, Su, Dv, Sv);
            aff_normal_m(a.c, a.rx0 + (float)e * a.c.k0, a.ry0 + (float)e * a.c.k3, Du, Su, Dv, Sv, fh * edge_f(j, W), m);
            aff_normalize(m, n);
            n0[e] = n[0]; n1[e] = n[1]; n2[e] = n[2];
        }
    }
    float* o = out + (long)q.pl * 3 * plane;
    store4(o, W, q.i, q.j, n0, vout); store4(o + plane, W, q.i, q.j, n1, vout); store4(o + 2 * plane, W, q.i, q.j, n2, vout);
}
// adjoints (R, L, Dn, Up) of the 4 pixels (i, j .. j+3) -> four shared planes (zeros outside the image)
__device__ __forceinline__ void aff_adj_quad(const float* __restrict__ p, const float* __restrict__ gp, long plane, AffPlane& a,
                                             const double* __restrict__ cam, int H, int W, int i, int j, bool vin,
                                             float* __restrict__ sR, float* __restrict__ sL, float* __restrict__ sD, float* __restrict__ sU) {
    float R[4] = {0.f, 0.f, 0.f, 0.f}, L[4] = {0.f, 0.f, 0.f, 0.f}, Dn[4] = {0.f, 0.f, 0.f, 0.f}, Up[4] = {0.f, 0.f, 0.f, 0.f};
    if (i >= 0 && i < H && j < W) {
        float u[4], l[4], c[6], g0[4], g1[4], g2[4];
        load4(p + (long)max(i - 1, 0) * W, j, W, vin, u);
        load4(p + (long)min(i + 1, H - 1) * W, j, W, vin, l);
        load6(p + (long)i * W, j, W, vin, c);
        load4(gp + (long)i * W, j, W, vin, g0);
        load4(gp + plane + (long)i * W, j, W, vin, g1);
        load4(gp + 2 * plane + (long)i * W, j, W, vin, g2);
        a.at(cam, i, j);
        if (i > 0 && i < H - 1 && j > 0 && j + 4 < W) {          // no image border in reach: literal masks / factors
#pragma unroll
            for (int e = 0; e < 4; ++e)
                aff_pixel_adj(a.c, a.rx0 + (float)e * a.c.k0, a.ry0 + (float)e * a.c.k3, c[e], c[e + 2], u[e], l[e],
                              1.f, 1.f, 1.f, 1.f, 0.25f, g0[e], g1[e], g2[e], R[e], L[e], Dn[e], Up[e]);
        } else {
            const float mu = i > 0 ? 1.f : 0.f, md = i < H - 1 ? 1.f : 0.f, fh = edge_f(i, H);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (j + e < W)
                    aff_pixel_adj(a.c, a.rx0 + (float)e * a.c.k0, a.ry0 + (float)e * a.c.k3, c[e], c[e + 2], u[e], l[e],
                                  j + e > 0 ? 1.f : 0.f, j + e < W - 1 ? 1.f : 0.f, mu, md, fh * edge_f(j + e, W), g0[e], g1[e], g2[e],
                                  R[e], L[e], Dn[e], Up[e]);
        }
    }
    *reinterpret_cast<float4*>(sR) = make_float4(R[0], R[1], R[2], R[3]);
    *reinterpret_cast<float4*>(sL) = make_float4(L[0], L[1], L[2], L[3]);
    *reinterpret_cast<float4*>(sD) = make_float4(Dn[0], Dn[1], Dn[2], Dn[3]);
    *reinterpret_cast<float4*>(sU) = make_float4(Up[0], Up[1], Up[2], Up[3]);
}
// same tiling as normals_old_bwd_quad: own quad per thread, 64 threads add the halo; shared planes: row r <-> i0 - 1 + r,
// column c <-> j0 - 4 + c.  gd(i, j) = R(i, j-1) + L(i, j+1) + Dn(i-1, j) + Up(i+1, j) + the pixel's own term where its
// neighbour was border-clamped onto itself.
__device__ __noinline__ float normals_generic_bwd_px(const float* __restrict__ p, const float* __restrict__ gp, long plane,
                                                     const double* __restrict__ cam, int H, int W, int i, int j) {
    return new_normal_bwd(p, gp, plane, cam, H, W, i, j);
}
// 4 CTAs per SM (64 registers; the fast path fits, the rare fp64 path spills): measured r99 252 us against 304 us at 3 CTAs
// (80 registers) and 407 us at 2 (118) on 96 x 512 x 640 - the kernel waits on its barrier and on loads, so occupancy pays
__global__ void __launch_bounds__(NT, 4)
normals_new_bwd_quad(const float* __restrict__ d, const float* __restrict__ g, const double* __restrict__ cams, int H, int W,
                     float* __restrict__ gd) {
    __shared__ __align__(16) float aR[(TH + 2) * AW], aL[(TH + 2) * AW], aD[(TH + 2) * AW], aU[(TH + 2) * AW];
    const double* cam = cams + blockIdx.z * DSR_CAM_DOUBLES;
    AffPlane a(cam);
    const int i0 = blockIdx.y * TH, j0 = blockIdx.x * TW;
    const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) << 2;
    const long plane = (long)H * W;
    const float* p = d + blockIdx.z * plane;
    const float* gp = g + (long)blockIdx.z * 3 * plane;
    if (!a.affine) {                                          // block-uniform: the whole plane takes the per-pixel fp64 form
        const int i = i0 + ty;
        if (i < H)
            for (int e = 0; e < 4 && j0 + tx + e < W; ++e)
                gd[blockIdx.z * plane + (long)i * W + j0 + tx + e] = normals_generic_bwd_px(p, gp, plane, cam, H, W, i, j0 + tx + e);
        return;
    }
    const bool vin = vec_ok(d, W) && vec_ok(g, W), vout = vec_ok(gd, W);
    {
        const int o = (ty + 1) * AW + 4 + tx;
        aff_adj_quad(p, gp, plane, a, cam, H, W, i0 + ty, j0 + tx, vin, aR + o, aL + o, aD + o, aU + o);
    }
    if (threadIdx.x < 32) {                       // rows i0 - 1 and i0 + TH
        const int r = (threadIdx.x >> 4) ? TH + 1 : 0, o = r * AW + 4 + tx;
        aff_adj_quad(p, gp, plane, a, cam, H, W, i0 - 1 + r, j0 + tx, vin, aR + o, aL + o, aD + o, aU + o);
    } else if (threadIdx.x < 64) {                // columns j0 - 1 (its R is read) and j0 + TW (its L) of the tile's rows
        const int t = threadIdx.x - 32, r = 1 + (t & 15), c = (t >> 4) ? TW + 4 : 3;
        const int i = i0 - 1 + r, j = j0 - 4 + c;
        float R = 0.f, L = 0.f, Dn = 0.f, Up = 0.f;
        if (i < H && j >= 0 && j < W) {
            const float* row = p + (long)i * W;
            const long o = (long)i * W + j;
            a.at(cam, i, j);
            aff_pixel_adj(a.c, a.rx0, a.ry0, __ldg(row + max(j - 1, 0)), __ldg(row + min(j + 1, W - 1)),
                          __ldg(p + (long)max(i - 1, 0) * W + j), __ldg(p + (long)min(i + 1, H - 1) * W + j),
                          j > 0 ? 1.f : 0.f, j < W - 1 ? 1.f : 0.f, i > 0 ? 1.f : 0.f, i < H - 1 ? 1.f : 0.f,
                          edge_f(i, H) * edge_f(j, W), __ldg(gp + o), __ldg(gp + plane + o), __ldg(gp + 2 * plane + o), R, L, Dn, Up);
        }
        aR[r * AW + c] = R; aL[r * AW + c] = L;
    }
    __syncthreads();
    const int i = i0 + ty, j = j0 + tx;
    if (i >= H || j >= W) return;
    const int o = (ty + 1) * AW + 4 + tx;
    const float4 r4 = *reinterpret_cast<const float4*>(aR + o), l4 = *reinterpret_cast<const float4*>(aL + o),
                 dn = *reinterpret_cast<const float4*>(aD + o - AW), up = *reinterpret_cast<const float4*>(aU + o + AW);
    const float Rv[5] = {aR[o - 1], r4.x, r4.y, r4.z, r4.w}, Lv[5] = {l4.x, l4.y, l4.z, l4.w, aL[o + 4]};
    const float Dv[4] = {dn.x, dn.y, dn.z, dn.w}, Uv[4] = {up.x, up.y, up.z, up.w};
    float self_d[4] = {0.f, 0.f, 0.f, 0.f};
    if (i == H - 1) { const float4 t = *reinterpret_cast<const float4*>(aD + o); self_d[0] += t.x; self_d[1] += t.y; self_d[2] += t.z; self_d[3] += t.w; }
    if (i == 0) { const float4 t = *reinterpret_cast<const float4*>(aU + o); self_d[0] += t.x; self_d[1] += t.y; self_d[2] += t.z; self_d[3] += t.w; }
    float out4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float v = Rv[e] + Lv[e + 1] + Dv[e] + Uv[e] + self_d[e];
        if (j + e == W - 1) v += Rv[e + 1];
        if (j + e == 0) v += Lv[e];
        out4[e] = v;
    }
    store4(gd + blockIdx.z * plane, W, i, j, out4, vout);
}

// ------------------------------------------------------------------------------------------
// total variation (main_model.py:15-19): sum of squared forward differences over (planes, H, W)
// ------------------------------------------------------------------------------------------
#define STRIP 64                             // reduction kernels: one CTA = 64 rows x 64 columns, 4 rows per thread
__global__ void __launch_bounds__(NT)
tv_fwd_quad(const float* __restrict__ x, int H, int W, double* __restrict__ out) {
    __shared__ double red[32];
    const int j = blockIdx.x * TW + ((threadIdx.x & 15) << 2);
    const int ibase = blockIdx.y * STRIP + (threadIdx.x >> 4) * 4;
    const float* p = x + (long)blockIdx.z * H * W;
    const bool vin = vec_ok(x, W);
    float a = 0.f;
    if (j < W && ibase < H) {
        float c[6], dn[4];
        load6(p + (long)ibase * W, j, W, vin, c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = ibase + k;
            if (i >= H) break;
            load4(p + (long)min(i + 1, H - 1) * W, j, W, vin, dn);     // clamped: zero difference past the border
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (j + e < W) {
                    const float r = c[e + 1] - ((j + e < W - 1) ? c[e + 2] : c[e + 1]), v = c[e + 1] - dn[e];
                    a += r * r + v * v;
                }
            }
            if (k < 3 && i + 1 < H) load6(p + (long)(i + 1) * W, j, W, vin, c);
        }
    }
    const double acc = block_sum<double>((double)a, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}
__global__ void __launch_bounds__(NT)
tv_bwd_quad(const float* __restrict__ x, int H, int W, const float* __restrict__ gscale, float coef, float* __restrict__ gx) {
    const Quad q(H, W);
    if (!q.ok) return;
    const long base = (long)q.pl * H * W;
    const bool vin = vec_ok(x, W), vout = vec_ok(gx, W);
    const float gs = coef * (gscale ? *gscale : 1.f) * 2.f;
    float u[4], l[4], c[6];
    load4(x + base + (long)max(q.i - 1, 0) * W, q.j, W, vin, u);       // clamped neighbours contribute exact zeros
    load4(x + base + (long)min(q.i + 1, H - 1) * W, q.j, W, vin, l);
    load6(x + base + (long)q.i * W, q.j, W, vin, c);
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float v = c[e + 1];
        const float right = (q.j + e < W - 1) ? c[e + 2] : v;          // (the !vec path clamps at W - 1 already)
        float a = v - right;
        a -= c[e] - v;
        a += v - l[e];
        a -= u[e] - v;
        o[e] = gs * a;
    }
    store4(gx + base, W, q.i, q.j, o, vout);
}

// ------------------------------------------------------------------------------------------
// masked L1 / L2:  t = (a*m1)*m2 - (b*m1)*m2 ; out[0] += sum|t| ; out[1] += sum t^2.  a, b (B, C, plane); masks
// (B, 1, plane), m2 may be null.  main_model.py:352,371-372,383-398.  grid.y walks the (b, c) planes.
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(NT)
masked_diff_fwd_v(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ m1,
                  const float* __restrict__ m2, int BC, int C, long plane, double* __restrict__ out) {
    __shared__ double red[32];
    double s1 = 0.0, s2 = 0.0;
    constexpr int V = VEC ? 4 : 1;
    const long nv = plane / V;
    for (int bc = blockIdx.y; bc < BC; bc += gridDim.y) {
        const float* pa = a + (long)bc * plane;
        const float* pb = b + (long)bc * plane;
        const float* pm = m1 + (long)(bc / C) * plane;
        const float* pn = m2 ? m2 + (long)(bc / C) * plane : nullptr;
        float f1 = 0.f, f2 = 0.f;
        for (long k = (long)blockIdx.x * NT + threadIdx.x; k < nv; k += (long)gridDim.x * NT) {
            float av[4], bv[4], mv[4];
            if (VEC) {
                const float4 x = ld4(pa + 4 * k), y = ld4(pb + 4 * k), m = ld4(pm + 4 * k);
                av[0] = x.x; av[1] = x.y; av[2] = x.z; av[3] = x.w; bv[0] = y.x; bv[1] = y.y; bv[2] = y.z; bv[3] = y.w;
                mv[0] = m.x; mv[1] = m.y; mv[2] = m.z; mv[3] = m.w;
            } else { av[0] = pa[k]; bv[0] = pb[k]; mv[0] = pm[k]; }
            float nv4[4] = {1.f, 1.f, 1.f, 1.f};
            if (pn) {
                if (VEC) { const float4 n = ld4(pn + 4 * k); nv4[0] = n.x; nv4[1] = n.y; nv4[2] = n.z; nv4[3] = n.w; }
                else nv4[0] = pn[k];
            }
#pragma unroll
            for (int e = 0; e < V; ++e) {
                float ta = av[e] * mv[e], tb = bv[e] * mv[e];
                if (pn) { ta *= nv4[e]; tb *= nv4[e]; }
                const float d = ta - tb;
                f1 += fabsf(d); f2 += d * d;
            }
        }
        s1 += (double)f1; s2 += (double)f2;
    }
    s1 = block_sum<double>(s1, red);
    s2 = block_sum<double>(s2, red);
    if (threadIdx.x == 0) { atomicAdd(out, s1); atomicAdd(out + 1, s2); }
}
// grad wrt b:  gb = -(c1*g1*sign(t) + c2*g2*2t) * m
template <bool VEC>
__global__ void __launch_bounds__(NT)
masked_diff_bwd_v(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ m1,
                  const float* __restrict__ m2, int BC, int C, long plane, const float* __restrict__ g1,
                  const float* __restrict__ g2, float c1, float c2, float* __restrict__ gb) {
    const float w1 = c1 * (g1 ? *g1 : 0.f), w2 = c2 * (g2 ? *g2 : 0.f);
    constexpr int V = VEC ? 4 : 1;
    const long nv = plane / V;
    for (int bc = blockIdx.y; bc < BC; bc += gridDim.y) {
        const float* pa = a + (long)bc * plane;
        const float* pb = b + (long)bc * plane;
        const float* pm = m1 + (long)(bc / C) * plane;
        const float* pn = m2 ? m2 + (long)(bc / C) * plane : nullptr;
        float* po = gb + (long)bc * plane;
        for (long k = (long)blockIdx.x * NT + threadIdx.x; k < nv; k += (long)gridDim.x * NT) {
            float av[4], bv[4], mv[4], ov[4];
            if (VEC) {
                const float4 x = ld4(pa + 4 * k), y = ld4(pb + 4 * k), m = ld4(pm + 4 * k);
                av[0] = x.x; av[1] = x.y; av[2] = x.z; av[3] = x.w; bv[0] = y.x; bv[1] = y.y; bv[2] = y.z; bv[3] = y.w;
                mv[0] = m.x; mv[1] = m.y; mv[2] = m.z; mv[3] = m.w;
            } else { av[0] = pa[k]; bv[0] = pb[k]; mv[0] = pm[k]; }
            float nv4[4] = {1.f, 1.f, 1.f, 1.f};
            if (pn) {
                if (VEC) { const float4 n = ld4(pn + 4 * k); nv4[0] = n.x; nv4[1] = n.y; nv4[2] = n.z; nv4[3] = n.w; }
                else nv4[0] = pn[k];
            }
#pragma unroll
            for (int e = 0; e < V; ++e) {
                float m = mv[e];
                float ta = av[e] * m, tb = bv[e] * m;
                if (pn) { ta *= nv4[e]; tb *= nv4[e]; m *= nv4[e]; }
                const float d = ta - tb;
                const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
                ov[e] = -(w1 * sg + w2 * 2.f * d) * m;
            }
            if (VEC) st4(po + 4 * k, make_float4(ov[0], ov[1], ov[2], ov[3])); else po[k] = ov[0];
        }
    }
}

// ------------------------------------------------------------------------------------------
// edge-aware smoothness, one pyramid level (main_model.py:22-73): d (B,1,h,w), img (B,C,h,w), C <= 4.
// 'x' = difference along H, 'y' = along W;  w = exp(-mean_c |grad I|);  out[0] += sum|dx wx|, out[1] += sum|dy wy|.
// Clamped neighbour rows / columns make every difference across the border an exact zero, so there are no border
// branches: |0 * w| = 0 and sign(0) = 0.
// ------------------------------------------------------------------------------------------
#define SM_MAXC 4
__device__ __forceinline__ float sgnf(float s) { return (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f); }
__global__ void __launch_bounds__(NT, 4)
smooth_fwd_quad(const float* __restrict__ d, const float* __restrict__ img, int C, int h, int w, double* __restrict__ out) {
    __shared__ double red[32];
    const int j = blockIdx.x * TW + ((threadIdx.x & 15) << 2);
    const int ibase = blockIdx.y * STRIP + (threadIdx.x >> 4) * 4;
    const int b = blockIdx.z;
    const long plane = (long)h * w;
    const float* pd = d + b * plane;
    const float* pi = img + (long)b * C * plane;
    const bool vin = vec_ok(d, w) && vec_ok(img, w);
    const float invC = 1.f / (float)C;
    float fx = 0.f, fy = 0.f;
    if (j < w) {
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int i = ibase + k;
            if (i >= h) break;
            const long r0 = (long)i * w, r1 = (long)min(i + 1, h - 1) * w;
            float c[6], l[4], sx[4] = {0.f, 0.f, 0.f, 0.f}, sy[4] = {0.f, 0.f, 0.f, 0.f};
            load6(pd + r0, j, w, vin, c);
            load4(pd + r1, j, w, vin, l);
            for (int ch = 0; ch < C; ++ch) {
                float ic[6], il[4];
                load6(pi + ch * plane + r0, j, w, vin, ic);
                load4(pi + ch * plane + r1, j, w, vin, il);
#pragma unroll
                for (int e = 0; e < 4; ++e) { sx[e] += fabsf(ic[e + 1] - il[e]); sy[e] += fabsf(ic[e + 1] - ic[e + 2]); }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (j + e < w) {
                    fx += fabsf((c[e + 1] - l[e]) * __expf(-sx[e] * invC));
                    fy += fabsf((c[e + 1] - c[e + 2]) * __expf(-sy[e] * invC));
                }
            }
        }
    }
    const double ax = block_sum<double>((double)fx, red), ay = block_sum<double>((double)fy, red);
    if (threadIdx.x == 0) { atomicAdd(out, ax); atomicAdd(out + 1, ay); }
}
// gd (=|+=) g * (cx * d/dd sum|dx wx| + cy * d/dd sum|dy wy|)
__global__ void __launch_bounds__(NT, 4)
smooth_bwd_quad(const float* __restrict__ d, const float* __restrict__ img, int C, int h, int w,
                const float* __restrict__ gscale, float cx, float cy, float* __restrict__ gd, int accumulate) {
    const Quad q(h, w);
    if (!q.ok) return;
    const long plane = (long)h * w;
    const float* pd = d + q.pl * plane;
    const float* pi = img + (long)q.pl * C * plane;
    const bool vin = vec_ok(d, w) && vec_ok(img, w), vout = vec_ok(gd, w);
    const float g = gscale ? *gscale : 1.f, invC = 1.f / (float)C;
    const long ru = (long)max(q.i - 1, 0) * w, rc = (long)q.i * w, rl = (long)min(q.i + 1, h - 1) * w;
    float u[4], c[6], l[4];
    load4(pd + ru, q.j, w, vin, u);
    load6(pd + rc, q.j, w, vin, c);
    load4(pd + rl, q.j, w, vin, l);
    float su[4] = {0.f, 0.f, 0.f, 0.f}, sl[4] = {0.f, 0.f, 0.f, 0.f}, sh[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ch = 0; ch < C; ++ch) {
        float iu[4], ic[6], il[4];
        load4(pi + ch * plane + ru, q.j, w, vin, iu);
        load6(pi + ch * plane + rc, q.j, w, vin, ic);
        load4(pi + ch * plane + rl, q.j, w, vin, il);
#pragma unroll
        for (int e = 0; e < 4; ++e) { su[e] += fabsf(iu[e] - ic[e + 1]); sl[e] += fabsf(ic[e + 1] - il[e]); }
#pragma unroll
        for (int e = 0; e < 5; ++e) sh[e] += fabsf(ic[e] - ic[e + 1]);
    }
    float wh[5];
#pragma unroll
    for (int e = 0; e < 5; ++e) wh[e] = __expf(-sh[e] * invC);          // weight of the (j+e-1, j+e) column pair
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float v = c[e + 1];
        const float wl = __expf(-sl[e] * invC), wu = __expf(-su[e] * invC);
        float acc = cx * wl * sgnf((v - l[e]) * wl);
        acc -= cx * wu * sgnf((u[e] - v) * wu);
        acc += cy * wh[e + 1] * sgnf((v - c[e + 2]) * wh[e + 1]);
        acc -= cy * wh[e] * sgnf((c[e] - v) * wh[e]);
        o[e] = g * acc;
    }
    float* row = gd + q.pl * plane;
    if (accumulate) {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (q.j + e < w) row[rc + q.j + e] += o[e];
    } else {
        store4(row, w, q.i, q.j, o, vout);
    }
}

// ---- rolling-row forms (16-byte aligned rows): one thread = 4 columns x 4 rows.  Every row of every plane is loaded ONCE per
// thread as a float4 (5 rows forward, 6 rows backward, all loads of a plane issued back to back), the row below / above is the
// register copy of the neighbouring row, and the column neighbours j-1 / j+4 come from the adjacent lanes by warp shuffle
// (a scalar load only at the 64-column tile seam).  The first-generation quad kernels loaded every row two (forward) or
// three (backward) times plus two scalar edge loads per row and plane, and ran at 31-37 % of the HBM peak, bound by load
// latency (ncu: long-scoreboard stalls 10 per issue, 36 % occupancy).
__device__ __forceinline__ float right_of(const float4 v, const float* __restrict__ row, int j, int w, int tx) {
    float nx = __shfl_down_sync(0xffffffffu, v.x, 1);              // x[j + 4] lives in the next lane's .x
    if (j + 4 >= w) nx = v.w;                                       // clamped: the difference across the border is 0
    else if (tx == 15) nx = __ldg(row + j + 4);                     // tile seam
    return nx;
}
__device__ __forceinline__ float left_of(const float4 v, const float* __restrict__ row, int j, int w, int tx) {
    float px = __shfl_up_sync(0xffffffffu, v.w, 1);                // x[j - 1] lives in the previous lane's .w
    if (j == 0 || j >= w) px = v.x;
    else if (tx == 0) px = __ldg(row + j - 1);
    return px;
}
template <int C>
__global__ void __launch_bounds__(NT, 2)
smooth_fwd_roll(const float* __restrict__ d, const float* __restrict__ img, int h, int w, double* __restrict__ out) {
    __shared__ double red[32];
    const int tx = threadIdx.x & 15;
    const int j = blockIdx.x * TW + (tx << 2);
    const int ibase = blockIdx.y * STRIP + (threadIdx.x >> 4) * 4;
    const int b = blockIdx.z;
    const long plane = (long)h * w;
    const int jc = j < w ? j : 0;                                   // lanes right of the image load column 0 and contribute 0
    float sx[4][4], sy[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) { sx[k][e] = 0.f; sy[k][e] = 0.f; }
    long roff[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) roff[k] = (long)min(ibase + k, h - 1) * w;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        const float* p = img + ((long)b * C + ch) * plane;
        float4 r[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) r[k] = ld4(p + roff[k] + jc);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float nx = right_of(r[k], p + roff[k], j, w, tx);
            sx[k][0] += fabsf(r[k].x - r[k + 1].x); sx[k][1] += fabsf(r[k].y - r[k + 1].y);
            sx[k][2] += fabsf(r[k].z - r[k + 1].z); sx[k][3] += fabsf(r[k].w - r[k + 1].w);
            sy[k][0] += fabsf(r[k].x - r[k].y); sy[k][1] += fabsf(r[k].y - r[k].z);
            sy[k][2] += fabsf(r[k].z - r[k].w); sy[k][3] += fabsf(r[k].w - nx);
        }
    }
    float fx = 0.f, fy = 0.f;
    {
        const float invC = 1.f / (float)C;
        const float* p = d + b * plane;
        float4 r[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) r[k] = ld4(p + roff[k] + jc);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float nx = right_of(r[k], p + roff[k], j, w, tx);
            if (j < w && ibase + k < h) {
                const float c[5] = {r[k].x, r[k].y, r[k].z, r[k].w, nx}, l[4] = {r[k + 1].x, r[k + 1].y, r[k + 1].z, r[k + 1].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    fx += fabsf((c[e] - l[e]) * __expf(-sx[k][e] * invC));
                    fy += fabsf((c[e] - c[e + 1]) * __expf(-sy[k][e] * invC));
                }
            }
        }
    }
    const double ax = block_sum<double>((double)fx, red), ay = block_sum<double>((double)fy, red);
    if (threadIdx.x == 0) { atomicAdd(out, ax); atomicAdd(out + 1, ay); }
}

template <int C>
__global__ void __launch_bounds__(NT, 2)
smooth_bwd_roll(const float* __restrict__ d, const float* __restrict__ img, int h, int w, const float* __restrict__ gscale,
                float cx, float cy, float* __restrict__ gd, int accumulate) {
    const int tx = threadIdx.x & 15;
    const int j = blockIdx.x * TW + (tx << 2);
    const int ibase = blockIdx.y * STRIP + (threadIdx.x >> 4) * 4;
    const int b = blockIdx.z;
    const long plane = (long)h * w;
    const int jc = j < w ? j : 0;
    // sv[k][e]: sum_c |I[i-1+k] - I[i+k]| at column j+e (row pairs k = 0..4); sh[k][e]: sum_c |I[.][j+e-1] - I[.][j+e]| of row i+k
    float sv[5][4], sh[4][5];
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) sv[k][e] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 5; ++e) sh[k][e] = 0.f;
    long roff[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) roff[k] = (long)min(max(ibase - 1 + k, 0), h - 1) * w;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        const float* p = img + ((long)b * C + ch) * plane;
        float4 r[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) r[k] = ld4(p + roff[k] + jc);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            sv[k][0] += fabsf(r[k].x - r[k + 1].x); sv[k][1] += fabsf(r[k].y - r[k + 1].y);
            sv[k][2] += fabsf(r[k].z - r[k + 1].z); sv[k][3] += fabsf(r[k].w - r[k + 1].w);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 v = r[k + 1];
            const float px = left_of(v, p + roff[k + 1], j, w, tx), nx = right_of(v, p + roff[k + 1], j, w, tx);
            sh[k][0] += fabsf(px - v.x); sh[k][1] += fabsf(v.x - v.y); sh[k][2] += fabsf(v.y - v.z);
            sh[k][3] += fabsf(v.z - v.w); sh[k][4] += fabsf(v.w - nx);
        }
    }
    const float g = gscale ? *gscale : 1.f, invC = 1.f / (float)C;
    const float* p = d + b * plane;
    float4 r[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) r[k] = ld4(p + roff[k] + jc);
    float* orow = gd + b * plane;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 v4 = r[k + 1];
        const float px = left_of(v4, p + roff[k + 1], j, w, tx), nx = right_of(v4, p + roff[k + 1], j, w, tx);
        if (j >= w || ibase + k >= h) continue;
        const float c[6] = {px, v4.x, v4.y, v4.z, v4.w, nx};
        const float u[4] = {r[k].x, r[k].y, r[k].z, r[k].w}, l[4] = {r[k + 2].x, r[k + 2].y, r[k + 2].z, r[k + 2].w};
        float wh[5], o[4];
#pragma unroll
        for (int e = 0; e < 5; ++e) wh[e] = __expf(-sh[k][e] * invC);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float v = c[e + 1];
            const float wl = __expf(-sv[k + 1][e] * invC), wu = __expf(-sv[k][e] * invC);
            float acc = cx * wl * sgnf((v - l[e]) * wl);
            acc -= cx * wu * sgnf((u[e] - v) * wu);
            acc += cy * wh[e + 1] * sgnf((v - c[e + 2]) * wh[e + 1]);
            acc -= cy * wh[e] * sgnf((c[e] - v) * wh[e]);
            o[e] = g * acc;
        }
        float* op = orow + (long)(ibase + k) * w + j;
        if (accumulate) {
            const float4 old = ld4(op);
            st4(op, make_float4(old.x + o[0], old.y + o[1], old.z + o[2], old.w + o[3]));
        } else {
            st4(op, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
// warp-strip rolling-row forms (csrc/stencil_roll.cu): 1 = launched, 0 = shape does not suit (ragged rows), take the quads
int dsr_roll_normals_old_fwd(const float* d, int B, int H, int W, float scale, float* out, void* stream);
int dsr_roll_normals_old_bwd(const float* d, const float* g, int B, int H, int W, float scale, float* gd, void* stream);
int dsr_roll_normals_new_fwd(const float* d, const double* cams, int B, int H, int W, float* out, void* stream);
int dsr_roll_normals_new_bwd(const float* d, const float* g, const double* cams, int B, int H, int W, float* gd, void* stream);
int dsr_roll_tv_fwd(const float* x, long planes, int H, int W, double* out, void* stream);
int dsr_roll_smooth_fwd(const float* d, const float* img, int B, int C, int H, int W, double* out2, void* stream);
int dsr_roll_smooth_bwd(const float* d, const float* img, int B, int C, int H, int W, const float* gscale, float cx, float cy, float* gd,
                        void* stream);
static int smooth_mode() {             // DSR_SMOOTH_KERNEL: 0 = automatic (strip kernels), 1 = ring / rolling-row kernels of the first two generations
    static int v = -1;
    if (v < 0) { const char* e = getenv("DSR_SMOOTH_KERNEL"); v = e ? atoi(e) : 0; }
    return v;
}
static bool use_roll() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DSR_STENCIL_ROLL"); v = e ? atoi(e) : 1; }
    return v != 0;
}

#define PLANES_OK(B, H, W) ((long)(B) <= 65535 && (((H) + TH - 1) / TH) <= 65535 && (long)(B) * (((H) + TH - 1) / TH) * (((W) + TW - 1) / TW) < (1L << 31) && (long)(H) * (W) < (1L << 31))

extern "C" int dsr_hole_valid_masks(const float* depth, int B, int H, int W, float border, float* hole,
                                    float* valid, void* stream) {
    DSR_REQUIRE(depth && valid && B > 0 && H > 0 && W > 0 && PLANES_OK(B, H, W), "bad arguments");
    hole_valid_quad<<<quad_grid(B, H, W), NT, 0, ST(stream)>>>(depth, H, W, border, hole, valid);
    return dsr_check_launch("hole_valid_masks");
}

extern "C" int dsr_rect_holes(const float* valid, const float* depth, const int* rects, const int* counts,
                              int max_rects, int B, int H, int W, float extra_border, unsigned char* gt_mask,
                              float* masked, float* extra, void* stream) {
    DSR_REQUIRE(valid && depth && rects && counts && gt_mask && masked, "null pointer");
    DSR_REQUIRE(max_rects > 0 && max_rects <= 1024 && B > 0 && PLANES_OK(B, H, W), "max_rects / shape out of range");
    DSR_REQUIRE(((uintptr_t)rects & 15) == 0, "rectangle table must be 16-byte aligned");
    rect_holes_tiled<<<tile_grid(B, H, W), NT, max_rects * 4 * sizeof(int), ST(stream)>>>(
        valid, depth, rects, counts, max_rects, B, H, W, extra_border, gt_mask, masked, extra);
    return dsr_check_launch("rect_holes");
}

extern "C" int dsr_normals_old_fwd(const float* depth, int B, int H, int W, float scale, float* out, void* stream) {
    DSR_REQUIRE(depth && out && B > 0 && H >= 2 && W >= 2 && PLANES_OK(B, H, W), "bad arguments");
    if (use_roll() && dsr_roll_normals_old_fwd(depth, B, H, W, scale, out, stream)) return dsr_check_launch("normals_old_fwd");
    normals_old_fwd_quad<<<quad_grid(B, H, W), NT, 0, ST(stream)>>>(depth, H, W, scale, out);
    return dsr_check_launch("normals_old_fwd");
}
extern "C" int dsr_normals_old_bwd(const float* depth, const float* gout, int B, int H, int W, float scale,
                                   float* gdepth, void* stream) {
    DSR_REQUIRE(depth && gout && gdepth && B > 0 && H >= 2 && W >= 2 && PLANES_OK(B, H, W), "bad arguments");
    if (use_roll() && dsr_roll_normals_old_bwd(depth, gout, B, H, W, scale, gdepth, stream)) return dsr_check_launch("normals_old_bwd");
    normals_old_bwd_quad<<<quad_grid(B, H, W), NT, 0, ST(stream)>>>(depth, gout, H, W, scale, gdepth);
    return dsr_check_launch("normals_old_bwd");
}
extern "C" int dsr_normals_new_fwd(const float* depth, const double* cams, int B, int H, int W, float* out,
                                   void* stream) {
    DSR_REQUIRE(depth && cams && out && B > 0 && H >= 2 && W >= 2 && PLANES_OK(B, H, W), "bad arguments");
    if (use_roll() && dsr_roll_normals_new_fwd(depth, cams, B, H, W, out, stream)) return dsr_check_launch("normals_new_fwd");
    normals_new_fwd_quad<<<quad_grid(B, H, W), NT, 0, ST(stream)>>>(depth, cams, H, W, out);
    return dsr_check_launch("normals_new_fwd");
}
extern "C" int dsr_normals_new_bwd(const float* depth, const float* gout, const double* cams, int B, int H, int W,
                                   float* gdepth, void* stream) {
    DSR_REQUIRE(depth && gout && cams && gdepth && B > 0 && H >= 2 && W >= 2 && PLANES_OK(B, H, W), "bad arguments");
    if (use_roll() && dsr_roll_normals_new_bwd(depth, gout, cams, B, H, W, gdepth, stream)) return dsr_check_launch("normals_new_bwd");
    normals_new_bwd_quad<<<quad_grid(B, H, W), NT, 0, ST(stream)>>>(depth, gout, cams, H, W, gdepth);
    return dsr_check_launch("normals_new_bwd");
}

extern "C" int dsr_tv_fwd(const float* x, long planes, int H, int W, double* out_sum, void* stream) {
    DSR_REQUIRE(x && out_sum && planes > 0 && H > 0 && W > 0 && PLANES_OK(planes, H, W), "bad arguments");
    if (use_roll() && dsr_roll_tv_fwd(x, planes, H, W, out_sum, stream)) return dsr_check_launch("tv_fwd");
    tv_fwd_quad<<<dim3((W + TW - 1) / TW, (H + STRIP - 1) / STRIP, (unsigned)planes), NT, 0, ST(stream)>>>(x, H, W, out_sum);
    return dsr_check_launch("tv_fwd");
}
extern "C" int dsr_tv_bwd(const float* x, long planes, int H, int W, const float* gscale, float coef, float* gx,
                          void* stream) {
    DSR_REQUIRE(x && gx && planes > 0 && H > 0 && W > 0 && PLANES_OK(planes, H, W), "bad arguments");
    tv_bwd_quad<<<quad_grid((int)planes, H, W), NT, 0, ST(stream)>>>(x, H, W, gscale, coef, gx);
    return dsr_check_launch("tv_bwd");
}

static dim3 plane_grid(int BC, long plane, int V) {
    const long per = (plane / V + NT - 1) / NT;
    const long cap = (long)dsr_num_sms() * 8;
    long gy = BC < cap ? BC : cap;
    long gx = cap / gy; if (gx < 1) gx = 1; if (gx > per) gx = per; if (gx < 1) gx = 1;
    if (gy > 65535) gy = 65535;
    return dim3((unsigned)gx, (unsigned)gy);
}
extern "C" int dsr_masked_diff_fwd(const float* a, const float* b, const float* m1, const float* m2, int B, int C,
                                   long plane, double* out2, void* stream) {
    DSR_REQUIRE(a && b && m1 && out2 && B > 0 && C > 0 && plane > 0, "bad arguments");
    const bool v = (plane & 3) == 0 && !(((uintptr_t)a | (uintptr_t)b | (uintptr_t)m1 | (uintptr_t)m2) & 15);
    if (v) masked_diff_fwd_v<true><<<plane_grid(B * C, plane, 4), NT, 0, ST(stream)>>>(a, b, m1, m2, B * C, C, plane, out2);
    else masked_diff_fwd_v<false><<<plane_grid(B * C, plane, 1), NT, 0, ST(stream)>>>(a, b, m1, m2, B * C, C, plane, out2);
    return dsr_check_launch("masked_diff_fwd");
}
extern "C" int dsr_masked_diff_bwd(const float* a, const float* b, const float* m1, const float* m2, int B, int C,
                                   long plane, const float* g_l1, const float* g_l2, float c1, float c2, float* gb,
                                   void* stream) {
    DSR_REQUIRE(a && b && m1 && gb && B > 0 && C > 0 && plane > 0, "bad arguments");
    const bool v = (plane & 3) == 0 && !(((uintptr_t)a | (uintptr_t)b | (uintptr_t)m1 | (uintptr_t)m2 | (uintptr_t)gb) & 15);
    if (v) masked_diff_bwd_v<true><<<plane_grid(B * C, plane, 4), NT, 0, ST(stream)>>>(a, b, m1, m2, B * C, C, plane, g_l1, g_l2, c1, c2, gb);
    else masked_diff_bwd_v<false><<<plane_grid(B * C, plane, 1), NT, 0, ST(stream)>>>(a, b, m1, m2, B * C, C, plane, g_l1, g_l2, c1, c2, gb);
    return dsr_check_launch("masked_diff_bwd");
}

extern "C" int dsr_smooth_level_fwd(const float* d, const float* img, int B, int C, int h, int w, double* out2,
                                    void* stream) {
    DSR_REQUIRE(d && img && out2 && B > 0 && C >= 1 && C <= SM_MAXC && h > 0 && w > 0 && PLANES_OK(B, h, w), "bad arguments (C <= 4)");
    const dim3 grid((w + TW - 1) / TW, (h + STRIP - 1) / STRIP, B);
    if (use_roll() && smooth_mode() == 0 && dsr_roll_smooth_fwd(d, img, B, C, h, w, out2, stream)) return dsr_check_launch("smooth_level_fwd");
    if (!((uintptr_t)d & 15) && !((uintptr_t)img & 15) && dsr_smooth_ring_suits(B, C, h, w))
        return dsr_smooth_level_fwd_ring(d, img, B, C, h, w, out2, stream);          // large plane sets: csrc/stencil_ring.cu
    if (!(w & 3) && !((uintptr_t)d & 15) && !((uintptr_t)img & 15)) {
        switch (C) {
            case 1: smooth_fwd_roll<1><<<grid, NT, 0, ST(stream)>>>(d, img, h, w, out2); break;
            case 2: smooth_fwd_roll<2><<<grid, NT, 0, ST(stream)>>>(d, img, h, w, out2); break;
            case 3: smooth_fwd_roll<3><<<grid, NT, 0, ST(stream)>>>(d, img, h, w, out2); break;
            default: smooth_fwd_roll<4><<<grid, NT, 0, ST(stream)>>>(d, img, h, w, out2); break;
        }
    } else {
        smooth_fwd_quad<<<grid, NT, 0, ST(stream)>>>(d, img, C, h, w, out2);
    }
    return dsr_check_launch("smooth_level_fwd");
}
extern "C" int dsr_smooth_level_bwd(const float* d, const float* img, int B, int C, int h, int w, const float* gscale,
                                    float cx, float cy, float* gd, int accumulate, void* stream) {
    DSR_REQUIRE(d && img && gd && B > 0 && C >= 1 && C <= SM_MAXC && h > 0 && w > 0 && PLANES_OK(B, h, w), "bad arguments (C <= 4)");
    if (use_roll() && smooth_mode() == 0 && !accumulate && dsr_roll_smooth_bwd(d, img, B, C, h, w, gscale, cx, cy, gd, stream))
        return dsr_check_launch("smooth_level_bwd");
    if (!((uintptr_t)d & 15) && !((uintptr_t)img & 15) && !((uintptr_t)gd & 15) && dsr_smooth_ring_suits(B, C, h, w))
        return dsr_smooth_level_bwd_ring(d, img, B, C, h, w, gscale, cx, cy, gd, accumulate, stream);
    if (!(w & 3) && !((uintptr_t)d & 15) && !((uintptr_t)img & 15) && !((uintptr_t)gd & 15)) {
        const dim3 grid((w + TW - 1) / TW, (h + STRIP - 1) / STRIP, B);
        switch (C) {
            case 1: smooth_bwd_roll<1><<<grid, NT, 0, ST(stream)>>>(d, img, h, w, gscale, cx, cy, gd, accumulate); break;
            case 2: smooth_bwd_roll<2><<<grid, NT, 0, ST(stream)>>>(d, img, h, w, gscale, cx, cy, gd, accumulate); break;
            case 3: smooth_bwd_roll<3><<<grid, NT, 0, ST(stream)>>>(d, img, h, w, gscale, cx, cy, gd, accumulate); break;
            default: smooth_bwd_roll<4><<<grid, NT, 0, ST(stream)>>>(d, img, h, w, gscale, cx, cy, gd, accumulate); break;
        }
    } else {
        smooth_bwd_quad<<<quad_grid(B, h, w), NT, 0, ST(stream)>>>(d, img, C, h, w, gscale, cx, cy, gd, accumulate);
    }
    return dsr_check_launch("smooth_level_bwd");
}
