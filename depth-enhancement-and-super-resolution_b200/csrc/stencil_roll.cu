// Warp-strip rolling-row forms of the normal / TV stencils (NCHW fp32 planes, rows 16-byte aligned, W % 4 == 0).
// Reference call sites: models/norms.py:185-235 (image-space normals), :75-108 (camera-space normals),
// models/main_model.py:15-19 (TV).
//
// Why a third form: the register-quad kernels of stencil_tiled.cu load three rows per output row (two of them L1 hits that
// still cost LSU slots and clamped index arithmetic), and the backward kernels exchange per-pixel adjoints through shared
// memory with halo tasks and a CTA barrier - ncu showed them issue-bound (normals_new bwd: ~250 instructions per pixel,
// 38 % of the HBM peak; normals_old bwd 53 %; normals_new fwd 55 %).  Here ONE WARP owns a strip of 128 columns (lane =
// 4 adjacent pixels) and walks down a chunk of rows:
//   * every input row is loaded exactly once per thread as one 16-byte load (the rows above / below are the register copies
//     of the previous / next iteration; the next iteration's loads are issued before this iteration's math);
//   * the column neighbours x[j-1] / x[j+4] come from the adjacent lanes by warp shuffle (one scalar load at a strip seam);
//   * backward: a pixel's adjoint is split into what it adds to its right / left / lower / upper neighbour (R, L, Dn, Up);
//     R / L travel between lanes by shuffle, Dn / Up between rows in registers - no shared memory, no barrier.  The two
//     pixels just outside the strip whose R / L the edge lanes need are recomputed by lanes 0 and 31 (1/4 more math on
//     that warp, no extra lanes);
//   * the B * H rows are cut into chunks so that all warp tasks fit ONE wave of resident warps (no tail wave).
#include <stdlib.h>
#include "common.cuh"
#include "stencil_math.cuh"
#include "../../include/dsr_b200.h"

#define ST(s) ((cudaStream_t)(s))
#define RW 4                         // warps per CTA
#define RNT (RW * 32)
#define RCOLS 128                    // columns per warp strip
#define FULL 0xffffffffu

struct RollPlan {
    int strips, chunks, rows, ntasks;
};
// chunks of rows such that planes x strips x chunks warp tasks fill k whole waves of `warps_per_sm` resident warps (smallest
// k), at least 8 rows per chunk (each chunk re-reads one halo row above and below)
static bool roll_plan(long planes, int H, int W, int warps_per_sm, RollPlan& p) {
    p.strips = (W + RCOLS - 1) / RCOLS;
    const long cols = planes * p.strips;
    const long slots = (long)dsr_num_sms() * warps_per_sm;
    long chunks = 1;
    for (int k = 1; k <= 64; ++k) {
        chunks = k * slots / cols;
        if (chunks >= 1) break;
    }
    if (chunks < 1) chunks = 1;
    long rows = (H + chunks - 1) / chunks;
    if (rows < 8) rows = 8;
    if (rows > H) rows = H;
    p.rows = (int)rows;
    p.chunks = (int)((H + rows - 1) / rows);
    const long nt = cols * p.chunks;
    if (nt >= (1L << 31) - RW) return false;
    p.ntasks = (int)nt;
    return true;
}
template <typename K>
static int resident_warps(K kernel, int smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, RNT, smem) != cudaSuccess || nb < 1) nb = 4;
    return nb * RW;
}

struct Task {
    int pl, i0, i1, j, lane;
    bool act;
};
__device__ __forceinline__ bool roll_task(const RollPlan& pl, int H, int W, Task& t) {
    const int id = blockIdx.x * RW + (threadIdx.x >> 5);
    if (id >= pl.ntasks) return false;
    int q = id;
    const int s = q % pl.strips; q /= pl.strips;
    const int ch = q % pl.chunks;
    t.pl = q / pl.chunks;
    t.i0 = ch * pl.rows;
    t.i1 = min(t.i0 + pl.rows, H);
    t.lane = threadIdx.x & 31;
    t.j = s * RCOLS + t.lane * 4;
    t.act = t.j < W;
    return true;
}
// row r (clamped into the image) at the lane's 4 columns; zeros for lanes right of the image
__device__ __forceinline__ float4 ldrow(const float* __restrict__ p, int r, int H, int W, const Task& t) {
    if (!t.act) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(reinterpret_cast<const float4*>(p + (long)min(max(r, 0), H - 1) * W + t.j));
}
// row r if it lies inside the image, zeros otherwise (gradient rows)
__device__ __forceinline__ float4 ldrow_in(const float* __restrict__ p, int r, int H, int W, const Task& t) {
    if (!t.act || r < 0 || r >= H) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(reinterpret_cast<const float4*>(p + (long)r * W + t.j));
}
// x[r][j-1] / x[r][j+4], clamped to the image (clamping makes every difference across the border an exact zero and encodes
// np.gradient's one-sided border differences).  Called by all 32 lanes.
__device__ __forceinline__ float left_nb(const float4 v, const float* __restrict__ rowp, const Task& t) {
    float x = __shfl_up_sync(FULL, v.w, 1);
    if (t.lane == 0) x = t.j > 0 ? __ldg(rowp + t.j - 1) : v.x;
    return x;
}
__device__ __forceinline__ float right_nb(const float4 v, const float* __restrict__ rowp, int W, const Task& t) {
    float x = __shfl_down_sync(FULL, v.x, 1);
    if (t.j + 4 >= W) x = v.w;
    else if (t.lane == 31) x = __ldg(rowp + t.j + 4);
    return x;
}
__device__ __forceinline__ float edge_half(int i, int n) { return (i == 0 || i == n - 1) ? 1.f : 0.5f; }

// ------------------------------------------------------------------------------------------
// per-pixel operators.  fwd(e, dl, dr, du, dd, masks, fh, fw) -> n[3];  adj(..., g) -> (R, L, Dn, Up)
// `e` = column offset of the pixel from the lane's first column (-1 / 4 for the seam pixels)
// ------------------------------------------------------------------------------------------
struct OldOp {                                   // image-space normals, norms.py:185-190
    float scale;
    __device__ __forceinline__ bool begin(int) { return true; }
    __device__ __forceinline__ void row_begin(int, int) {}
    __device__ __forceinline__ void row_next() {}
    __device__ __forceinline__ void fwd(int, float dl, float dr, float du, float dd, float, float, float, float, float fh, float fw,
                                        float n[3]) const {
        const float gh = (dd - du) * fh, gw = (dr - dl) * fw;
        const float r = sqrtf(gh * gh + gw * gw + 1.f);
        const float k = scale / (r + 1e-6f);
        n[0] = -gh * k; n[1] = -gw * k; n[2] = k;
    }
    __device__ __forceinline__ void adj(int, float dl, float dr, float du, float dd, float, float, float, float, float fh, float fw,
                                        float g0, float g1, float g2, float& R, float& L, float& Dn, float& Up) const {
        const float gh = (dd - du) * fh, gw = (dr - dl) * fw;
        const float r2 = gh * gh + gw * gw + 1.f;
        const float ir = rsqrtf(r2), r = r2 * ir;
        const float iden = __fdividef(1.f, r + 1e-6f);
        const float dn0 = g0 * scale, dn1 = g1 * scale, dn2 = g2 * scale;
        const float dot = dn2 - dn0 * gh - dn1 * gw;
        const float k = dot * ir * iden * iden;
        const float dgh = -(dn0 * iden + k * gh) * fh, dgw = -(dn1 * iden + k * gw) * fw;
        Dn = dgh; Up = -dgh; R = dgw; L = -dgw;
    }
};
struct NewOp {                                   // camera-space normals, closed fp32 form (stencil_math.cuh aff_*)
    const double* cams;
    const double* cam;
    AffCam c;
    float rx0, ry0;
    double rxd, ryd, k1d, k4d;                   // ray of the lane's first pixel in the current row, advanced in fp64 (exact to
                                                 // 1e-16: the fp32 copies equal the per-row evaluation up to a rounding tie)
    __device__ __forceinline__ bool begin(int pl) {
        cam = cams + (long)pl * DSR_CAM_DOUBLES;
        c = aff_cam(cam);
        return cam_is_affine(cam);
    }
    __device__ __forceinline__ void row_begin(int i, int j) {
        const double u = cam[9] + (double)j, v = cam[10] + (double)i;
        k1d = cam[1]; k4d = cam[4];
        rxd = cam[0] * u + k1d * v + cam[2];
        ryd = cam[3] * u + k4d * v + cam[5];
        rx0 = (float)rxd; ry0 = (float)ryd;
    }
    __device__ __forceinline__ void row_next() {
        rxd += k1d; ryd += k4d;
        rx0 = (float)rxd; ry0 = (float)ryd;
    }
    __device__ __forceinline__ void fwd(int e, float dl, float dr, float du, float dd, float ma, float mb, float mu, float md, float fh,
                                        float fw, float n[3]) const {
        float Du, Su, Dv, Sv, m[3];
        aff_terms(dl, dr, du, dd, ma, mb, mu, md, Du, Su, Dv, Sv);
        aff_normal_m(c, rx0 + (float)e * c.k0, ry0 + (float)e * c.k3, Du, Su, Dv, Sv, fh * fw, m);
        aff_normalize(m, n);
    }
    __device__ __forceinline__ void adj(int e, float dl, float dr, float du, float dd, float ma, float mb, float mu, float md, float fh,
                                        float fw, float g0, float g1, float g2, float& R, float& L, float& Dn, float& Up) const {
        aff_pixel_adj(c, rx0 + (float)e * c.k0, ry0 + (float)e * c.k3, dl, dr, du, dd, ma, mb, mu, md, fh * fw, g0, g1, g2, R, L, Dn, Up);
    }
};
__device__ __noinline__ void roll_generic_fwd(const float* __restrict__ p, const double* __restrict__ cam, int H, int W, const Task& t,
                                              float* __restrict__ o, long plane) {
    if (!t.act) return;
    for (int i = t.i0; i < t.i1; ++i)
        for (int e = 0; e < 4; ++e) {
            float n[3];
            new_normal_fwd(p, cam, H, W, i, t.j + e, n);
            const long at = (long)i * W + t.j + e;
            o[at] = n[0]; o[plane + at] = n[1]; o[2 * plane + at] = n[2];
        }
}
__device__ __noinline__ void roll_generic_bwd(const float* __restrict__ p, const float* __restrict__ gp, const double* __restrict__ cam,
                                              int H, int W, const Task& t, float* __restrict__ o, long plane) {
    if (!t.act) return;
    for (int i = t.i0; i < t.i1; ++i)
        for (int e = 0; e < 4; ++e) o[(long)i * W + t.j + e] = new_normal_bwd(p, gp, plane, cam, H, W, i, t.j + e);
}

// ------------------------------------------------------------------------------------------
// RowPipe: the next DEPTH rows of every input plane in flight WITHOUT holding registers - each thread copies its own 16 bytes
// per plane and row (plus the few scalars the strip-seam lanes need) into a private shared-memory slot with cp.async and
// reads it back DEPTH iterations later.  Only the issuing thread reads a slot, so cp.async.wait_group is the only
// synchronisation.  Measured before this existed (r2b): with one row of register prefetch the rolling kernels had ~40 KB of
// loads in flight per SM and ran at 27-65 % of the HBM peak - slower than the quads; the HBM queue needs >= 100 KB per SM.
// Shared layout per CTA: [DEPTH][NV float4 units][RNT] then [DEPTH][NS words][RNT] (lanes contiguous: conflict-free).
// ------------------------------------------------------------------------------------------
template <int NV, int NS, int DEPTH, int NT = RNT>
struct RowPipe {
    uint32_t vbase, sbase;
    static constexpr int bytes(int nt = NT) { return DEPTH * (NV * 16 + NS * 4) * nt; }
    int nt;                                                          // threads per CTA (NT == 0: blockDim.x)
    __device__ __forceinline__ RowPipe() {
        extern __shared__ __align__(16) unsigned char roll_smem[];
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(roll_smem);
        nt = NT ? NT : (int)blockDim.x;
        vbase = b + threadIdx.x * 16;
        sbase = b + DEPTH * NV * 16 * nt + threadIdx.x * 4;
    }
    __device__ __forceinline__ void cp16(int slot, int u, const float* src, bool valid) const {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vbase + (slot * NV + u) * (nt * 16)), "l"(src), "r"(valid ? 16 : 0));
    }
    __device__ __forceinline__ void cp4(int slot, int w, const float* src, bool valid) const {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sbase + (slot * NS + w) * (nt * 4)), "l"(src), "r"(valid ? 4 : 0));
    }
    __device__ __forceinline__ float4 v(int slot, int u) const {
        float4 r;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(vbase + (slot * NV + u) * (nt * 16)));
        return r;
    }
    __device__ __forceinline__ float s(int slot, int w) const {
        float r;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(sbase + (slot * NS + w) * (nt * 4)));
        return r;
    }
    __device__ __forceinline__ static void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    __device__ __forceinline__ static void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory"); }
};
__device__ __forceinline__ int clampi(int v, int n) { return min(max(v, 0), n - 1); }

// ------------------------------------------------------------------------------------------
// forward: depth (planes, H, W) -> normals (planes, 3, H, W).  Pipe row k = depth row k + 1 (+ its seam neighbour)
// ------------------------------------------------------------------------------------------
#define FWD_DEPTH 6
template <class OP, bool NEW, int MINB>
__global__ void __launch_bounds__(RNT, MINB)
normals_fwd_roll(const float* __restrict__ d, OP op, int H, int W, RollPlan pl, float* __restrict__ out) {
    Task t;
    if (!roll_task(pl, H, W, t)) return;
    const long plane = (long)H * W;
    const float* p = d + t.pl * plane;
    float* o = out + (long)t.pl * 3 * plane;
    if (!op.begin(t.pl)) {
        if constexpr (NEW) roll_generic_fwd(p, op.cam, H, W, t, o, plane);
        return;
    }
    const RowPipe<1, 1, FWD_DEPTH> pipe;
    const bool seamL = t.lane == 0 && t.j > 0, seamR = t.lane == 31 && t.j + 4 < W, seam = seamL || seamR;
    const int jn = seamL ? t.j - 1 : (seamR ? t.j + 4 : 0);          // the seam lane's neighbour column
    const int jc = t.act ? t.j : 0;
    auto issue = [&](int k, int slot) {                              // 32-bit offsets inside the plane (H * W < 2^31)
        if (k < t.i1) {
            const int od = clampi(k + 1, H) * W;
            pipe.cp16(slot, 0, p + (od + jc), t.act);
            if (seam) pipe.cp4(slot, 0, p + (od + jn), true);
        }
        pipe.commit();
    };
#pragma unroll
    for (int k = 0; k < FWD_DEPTH; ++k) issue(t.i0 + k, k);
    float4 dU = ldrow(p, t.i0 - 1, H, W, t), dC = ldrow(p, t.i0, H, W, t);
    float sC = seam ? __ldg(p + (long)t.i0 * W + jn) : 0.f;
    const bool colin = t.j > 0 && t.j + 4 < W;
    op.row_begin(t.i0, t.j);
    int slot = 0;
    for (int r = t.i0; r < t.i1; ++r) {
        pipe.wait();
        const float4 dD = pipe.v(slot, 0);
        const float sD = pipe.s(slot, 0);
        issue(r + FWD_DEPTH, slot);
        slot = slot + 1 == FWD_DEPTH ? 0 : slot + 1;
        float xl = __shfl_up_sync(FULL, dC.w, 1), xr = __shfl_down_sync(FULL, dC.x, 1);
        if (t.lane == 0) xl = seamL ? sC : dC.x;
        if (t.j + 4 >= W) xr = dC.w; else if (t.lane == 31) xr = sC;
        const float c[6] = {xl, dC.x, dC.y, dC.z, dC.w, xr};
        const float u[4] = {dU.x, dU.y, dU.z, dU.w}, l[4] = {dD.x, dD.y, dD.z, dD.w};
        float n0[4], n1[4], n2[4];
        if (colin && r > 0 && r < H - 1) {                           // no image border in reach: literal masks / factors
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float n[3];
                op.fwd(e, c[e], c[e + 2], u[e], l[e], 1.f, 1.f, 1.f, 1.f, 0.5f, 0.5f, n);
                n0[e] = n[0]; n1[e] = n[1]; n2[e] = n[2];
            }
        } else {
            const float mu = r > 0 ? 1.f : 0.f, md = r < H - 1 ? 1.f : 0.f, fh = edge_half(r, H);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = t.j + e;
                float n[3];
                op.fwd(e, c[e], c[e + 2], u[e], l[e], j > 0 ? 1.f : 0.f, j < W - 1 ? 1.f : 0.f, mu, md, fh, edge_half(j, W), n);
                n0[e] = n[0]; n1[e] = n[1]; n2[e] = n[2];
            }
        }
        if (t.act) {
            float* q = o + (long)r * W + t.j;
            st4(q, make_float4(n0[0], n0[1], n0[2], n0[3]));
            st4(q + plane, make_float4(n1[0], n1[1], n1[2], n1[3]));
            st4(q + 2 * plane, make_float4(n2[0], n2[1], n2[2], n2[3]));
        }
        dU = dC; dC = dD; sC = sD;
        op.row_next();
    }
}

// ------------------------------------------------------------------------------------------
// backward: (depth, dL/dn) -> dL/ddepth.  gd(i, j) = R(i, j-1) + L(i, j+1) + Dn(i-1, j) + Up(i+1, j) + the pixel's own term
// where its neighbour was border-clamped onto itself.  Pipe row k = depth row k + 1 and the three gradient rows k, plus for
// the seam lanes depth[k+1][js], depth[k+1][jo] (js = the pixel just outside the strip, jo = its outer neighbour) and g[k][js].
// ------------------------------------------------------------------------------------------
#define BWD_DEPTH 3
template <class OP, bool NEW, int MINB>
__global__ void __launch_bounds__(RNT, MINB)
normals_bwd_roll(const float* __restrict__ d, const float* __restrict__ g, OP op, int H, int W, RollPlan pl, float* __restrict__ gd) {
    Task t;
    if (!roll_task(pl, H, W, t)) return;
    const long plane = (long)H * W;
    const float* p = d + t.pl * plane;
    const float* gp = g + (long)t.pl * 3 * plane;
    float* o = gd + t.pl * plane;
    if (!op.begin(t.pl)) {
        if constexpr (NEW) roll_generic_bwd(p, gp, op.cam, H, W, t, o, plane);
        return;
    }
    const RowPipe<4, 5, BWD_DEPTH> pipe;
    const bool seamL = t.lane == 0 && t.j > 0, seamR = t.lane == 31 && t.j + 4 < W, seam = seamL || seamR;
    const int js = seamL ? t.j - 1 : (seamR ? t.j + 4 : 0), es = seamL ? -1 : 4;
    const int jo = seamL ? max(js - 1, 0) : min(js + 1, W - 1);
    const int jc = t.act ? t.j : 0;
    const bool colin = t.j > 0 && t.j + 4 < W;
    const int r0 = t.i0 - 1;
    const float* gp1 = gp + plane;
    const float* gp2 = gp1 + plane;
    auto issue = [&](int k, int slot) {                              // 32-bit offsets inside the plane (H * W < 2^31)
        if (k <= t.i1) {
            const int od = clampi(k + 1, H) * W, og = clampi(k, H) * W;
            const bool gin = (unsigned)k < (unsigned)H, ga = gin && t.act;
            pipe.cp16(slot, 0, p + (od + jc), t.act);
            pipe.cp16(slot, 1, gp + (og + jc), ga);
            pipe.cp16(slot, 2, gp1 + (og + jc), ga);
            pipe.cp16(slot, 3, gp2 + (og + jc), ga);
            if (seam) {
                pipe.cp4(slot, 0, p + (od + js), true);
                pipe.cp4(slot, 1, p + (od + jo), true);
                pipe.cp4(slot, 2, gp + (og + js), gin);
                pipe.cp4(slot, 3, gp1 + (og + js), gin);
                pipe.cp4(slot, 4, gp2 + (og + js), gin);
            }
        }
        pipe.commit();
    };
#pragma unroll
    for (int k = 0; k < BWD_DEPTH; ++k) issue(r0 + k, k);
    float4 dU = ldrow(p, r0 - 1, H, W, t), dC = ldrow(p, r0, H, W, t);
    float sU = 0.f, sC = 0.f, sO = 0.f;                                // seam column: rows r - 1, r at js; row r at jo
    if (seam) {
        sU = __ldg(p + (long)clampi(r0 - 1, H) * W + js);
        sC = __ldg(p + (long)clampi(r0, H) * W + js);
        sO = __ldg(p + (long)clampi(r0, H) * W + jo);
    }
    float Hp[4] = {0.f, 0.f, 0.f, 0.f}, DnP[4] = {0.f, 0.f, 0.f, 0.f}, DnPP[4] = {0.f, 0.f, 0.f, 0.f};
    op.row_begin(r0, t.j);
    int slot = 0;
    for (int r = r0; r <= t.i1; ++r) {
        pipe.wait();
        const float4 dD = pipe.v(slot, 0), G0 = pipe.v(slot, 1), G1 = pipe.v(slot, 2), G2 = pipe.v(slot, 3);
        const float sD = pipe.s(slot, 0), sOn = pipe.s(slot, 1), sg0 = pipe.s(slot, 2), sg1 = pipe.s(slot, 3), sg2 = pipe.s(slot, 4);
        issue(r + BWD_DEPTH, slot);
        slot = slot + 1 == BWD_DEPTH ? 0 : slot + 1;
        float R[4] = {0.f, 0.f, 0.f, 0.f}, L[4] = {0.f, 0.f, 0.f, 0.f}, Dn[4] = {0.f, 0.f, 0.f, 0.f}, Up[4] = {0.f, 0.f, 0.f, 0.f};
        float Rs = 0.f, Ls = 0.f;
        if (r >= 0 && r < H) {                                        // warp-uniform
            float xl = __shfl_up_sync(FULL, dC.w, 1), xr = __shfl_down_sync(FULL, dC.x, 1);
            if (t.lane == 0) xl = seamL ? sC : dC.x;
            if (t.j + 4 >= W) xr = dC.w; else if (t.lane == 31) xr = sC;
            const float c[6] = {xl, dC.x, dC.y, dC.z, dC.w, xr};
            const float u[4] = {dU.x, dU.y, dU.z, dU.w}, l[4] = {dD.x, dD.y, dD.z, dD.w};
            const float g0[4] = {G0.x, G0.y, G0.z, G0.w}, g1[4] = {G1.x, G1.y, G1.z, G1.w}, g2[4] = {G2.x, G2.y, G2.z, G2.w};
            const float mu = r > 0 ? 1.f : 0.f, md = r < H - 1 ? 1.f : 0.f, fh = edge_half(r, H);
            if (colin && r > 0 && r < H - 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    op.adj(e, c[e], c[e + 2], u[e], l[e], 1.f, 1.f, 1.f, 1.f, 0.5f, 0.5f, g0[e], g1[e], g2[e], R[e], L[e], Dn[e], Up[e]);
            } else if (t.act) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = t.j + e;
                    op.adj(e, c[e], c[e + 2], u[e], l[e], j > 0 ? 1.f : 0.f, j < W - 1 ? 1.f : 0.f, mu, md, fh, edge_half(j, W),
                           g0[e], g1[e], g2[e], R[e], L[e], Dn[e], Up[e]);
                }
            }
            if (seam) {                                               // the pixel just outside the strip (lanes 0 / 31)
                const float sl = seamL ? sO : dC.w, sr = seamL ? dC.x : sO;
                float a, b, cc, dd;
                op.adj(es, sl, sr, sU, sD, js > 0 ? 1.f : 0.f, js < W - 1 ? 1.f : 0.f, mu, md, fh, edge_half(js, W), sg0, sg1, sg2,
                       a, b, cc, dd);
                Rs = seamL ? a : 0.f;
                Ls = seamR ? b : 0.f;
            }
        }
        // horizontal exchange (all lanes): R of the pixel on the left, L of the pixel on the right
        float Rl = __shfl_up_sync(FULL, R[3], 1), Lr = __shfl_down_sync(FULL, L[0], 1);
        if (t.lane == 0) Rl = Rs;
        if (t.lane == 31) Lr = Ls;
        float Hs[4] = {Rl + L[1], R[0] + L[2], R[1] + L[3], R[2] + Lr};
        if (t.j == 0) Hs[0] += L[0];
        if (t.j + 4 == W) Hs[3] += R[3];
        if (r == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) Hs[e] += Up[e];
        }
        if (r == H - 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) Hs[e] += Dn[e];
        }
        if (r > t.i0 && t.act)                                        // row r - 1 is complete
            st4(o + (long)(r - 1) * W + t.j,
                make_float4(Hp[0] + DnPP[0] + Up[0], Hp[1] + DnPP[1] + Up[1], Hp[2] + DnPP[2] + Up[2], Hp[3] + DnPP[3] + Up[3]));
#pragma unroll
        for (int e = 0; e < 4; ++e) { DnPP[e] = DnP[e]; DnP[e] = Dn[e]; Hp[e] = Hs[e]; }
        dU = dC; dC = dD; sU = sC; sC = sD; sO = sOn;
        op.row_next();
    }
}


// ------------------------------------------------------------------------------------------
// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot).  The stencil kernels are bound by
// instruction issue, not by the fp32 pipe (ncu r2d: 68 % issue-active at 27-48 % DRAM throughput), and every pixel of a
// thread's quad runs the same formula - so the quad is computed as two pairs.
// ------------------------------------------------------------------------------------------
struct f2 {
    float2 v;
};
__device__ __forceinline__ f2 mk(float a, float b) { f2 r; r.v = make_float2(a, b); return r; }
__device__ __forceinline__ f2 bc(float a) { return mk(a, a); }
__device__ __forceinline__ f2 mul(f2 a, f2 b) { f2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 add(f2 a, f2 b) { f2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 fma(f2 a, f2 b, f2 c) { f2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
__device__ __forceinline__ f2 sub(f2 a, f2 b) { return fma(b, bc(-1.f), a); }
__device__ __forceinline__ f2 inv_len(f2 r2) { return mk(rsqrtf(fmaxf(r2.v.x, 1e-24f)), rsqrtf(fmaxf(r2.v.y, 1e-24f))); }
__device__ __forceinline__ f2 keep_if_gt(f2 r2, float thr, f2 a) { return mk(r2.v.x > thr ? a.v.x : 0.f, r2.v.y > thr ? a.v.y : 0.f); }

// per-plane / per-row constants of the camera-space adjoint, as broadcast pairs
struct AffPair {
    f2 k0, k4, nk1, nk3, D;
    __device__ __forceinline__ void set(const AffCam& c) { k0 = bc(c.k0); k4 = bc(c.k4); nk1 = bc(-c.k1); nk3 = bc(-c.k3); D = bc(c.D); }
};
// Border handling without branches: hma / hmb = 0.5 * (left / right neighbour is a real pixel), hmab = hma + hmb, likewise
// hmu / hmd / hmud for the rows; sc = product of np.gradient's 1/2 (interior) or 1 (border) factors.
// m = Pv x Pu of stencil_math.cuh (aff_normal_m) for a pair of pixels
__device__ __forceinline__ void aff_m_pair(const AffPair& k, f2 nrx, f2 nry, f2 dl, f2 dr, f2 du, f2 dd, f2 hma, f2 hmb, f2 hmab, f2 hmu,
                                           f2 hmd, f2 hmud, f2 sc, f2& Du, f2& Su, f2& Dv, f2& Sv, f2& m0, f2& m1, f2& m2) {
    const f2 h = bc(0.5f);
    Du = mul(h, sub(dr, dl));
    Su = fma(hmb, dr, fma(hma, dl, hmab));
    Dv = mul(h, sub(dd, du));
    Sv = fma(hmd, dd, fma(hmu, du, hmud));
    const f2 A = mul(Su, Dv), B = mul(Sv, Du);
    m0 = mul(sc, fma(k.k4, B, mul(k.nk3, A)));
    m1 = mul(sc, fma(k.k0, A, mul(k.nk1, B)));
    m2 = fma(nrx, m0, fma(nry, m1, mul(mul(sc, k.D), mul(Su, Sv))));
}
// adjoint of a pair: dL/dn (g) -> what each pixel adds to dL/dd of its right / left / lower / upper neighbour
__device__ __forceinline__ void aff_adj_pair(const AffPair& k, f2 nrx, f2 nry, f2 dl, f2 dr, f2 du, f2 dd, f2 hma, f2 hmb, f2 hmab, f2 hmu,
                                             f2 hmd, f2 hmud, f2 sc, f2 g0, f2 g1, f2 g2, f2& R, f2& L, f2& Dn, f2& Up) {
    f2 Du, Su, Dv, Sv, m0, m1, m2;
    aff_m_pair(k, nrx, nry, dl, dr, du, dd, hma, hmb, hmab, hmu, hmd, hmud, sc, Du, Su, Dv, Sv, m0, m1, m2);
    const f2 r2 = fma(m0, m0, fma(m1, m1, mul(m2, m2)));
    const f2 ir = inv_len(r2);
    const f2 n0 = mul(m0, ir), n1 = mul(m1, ir), n2 = mul(m2, ir);
    const f2 dot = keep_if_gt(r2, 1e-24f, fma(n0, g0, fma(n1, g1, mul(n2, g2))));   // clamped denominator: n = m * 1e12, no projection
    const f2 ndot = mul(dot, bc(-1.f));
    const f2 dm0 = mul(fma(n0, ndot, g0), ir), dm1 = mul(fma(n1, ndot, g1), ir), dm2 = mul(fma(n2, ndot, g2), ir);
    const f2 e0 = fma(nrx, dm2, dm0), e1 = fma(nry, dm2, dm1), e2 = mul(mul(sc, k.D), dm2);
    const f2 p = mul(sc, fma(k.k4, e0, mul(k.nk1, e1))), q = mul(sc, fma(k.k0, e1, mul(k.nk3, e0)));
    const f2 dDu = mul(Sv, p), dSv = fma(Du, p, mul(e2, Su)), dDv = mul(Su, q), dSu = fma(Dv, q, mul(e2, Sv));
    const f2 hu = mul(bc(0.5f), dDu), hv = mul(bc(0.5f), dDv);
    R = fma(hmb, dSu, hu);
    L = sub(mul(hma, dSu), hu);
    Dn = fma(hmd, dSv, hv);
    Up = sub(mul(hmu, dSv), hv);
}
// image-space normals (norms.py:185-190): gh = (dd - du) fh, gw = (dr - dl) fw with the border factors folded into the pair
// constants fw2 = 2 * hm-style factors: fwp = fw (0.5 interior / 1 border) per pixel, fh broadcast
__device__ __forceinline__ void old_adj_pair(float scale, f2 dl, f2 dr, f2 du, f2 dd, f2 fh, f2 fw, f2 g0, f2 g1, f2 g2, f2& R, f2& L,
                                             f2& Dn, f2& Up) {
    const f2 gh = mul(sub(dd, du), fh), gw = mul(sub(dr, dl), fw);
    const f2 r2 = fma(gh, gh, fma(gw, gw, bc(1.f)));
    const f2 ir = mk(rsqrtf(r2.v.x), rsqrtf(r2.v.y));
    const f2 r = mul(r2, ir);
    const f2 den = add(r, bc(1e-6f));
    const f2 iden = mk(__fdividef(1.f, den.v.x), __fdividef(1.f, den.v.y));
    const f2 sc = bc(scale);
    const f2 dn0 = mul(g0, sc), dn1 = mul(g1, sc), dn2 = mul(g2, sc);
    const f2 ngh = mul(gh, bc(-1.f)), ngw = mul(gw, bc(-1.f));
    const f2 dot = fma(dn0, ngh, fma(dn1, ngw, dn2));
    const f2 kk = mul(mul(dot, ir), mul(iden, iden));
    // dgh = -(dn0 iden + k gh) fh ;  Dn = dgh, Up = -dgh ;  R = dgw, L = -dgw
    Up = mul(fma(dn0, iden, mul(kk, gh)), fh);
    Dn = mul(Up, bc(-1.f));
    L = mul(fma(dn1, iden, mul(kk, gw)), fw);
    R = mul(L, bc(-1.f));
}

struct OldBand {
    float scale;
    __device__ __forceinline__ bool begin(int) { return true; }
    __device__ __forceinline__ void row_begin(int, int) {}
    __device__ __forceinline__ void row_next() {}
    const double* cam;
};
struct NewBand {
    const double* cams;
    const double* cam;
    AffPair k;
    float k0;
    double rxd, ryd, k1d, k4d;
    float k3;
    __device__ __forceinline__ bool begin(int pl) {
        cam = cams + (long)pl * DSR_CAM_DOUBLES;
        const AffCam c = aff_cam(cam);
        k.set(c); k0 = c.k0; k3 = c.k3;
        return cam_is_affine(cam);
    }
    __device__ __forceinline__ void row_begin(int i, int j) {
        const double u = cam[9] + (double)j, v = cam[10] + (double)i;
        k1d = cam[1]; k4d = cam[4];
        rxd = cam[0] * u + k1d * v + cam[2];
        ryd = cam[3] * u + k4d * v + cam[5];
    }
    __device__ __forceinline__ void row_next() { rxd += k1d; ryd += k4d; }
};

// ------------------------------------------------------------------------------------------
// backward, row-band form: one CTA = every column of a band of rows (warp w = columns 128 w ..), so a contiguous piece of
// each plane streams through one CTA.  Per row: the adjoints' R / L and the next depth row's edge values cross the warp
// seams through 16 floats of shared memory and ONE barrier; nothing is recomputed and no lane idles.  W <= 1024.
// ------------------------------------------------------------------------------------------
#define BAND_DEPTH 3
#define BAND_MAXW 8
template <class OP, bool NEW, int MINB>
__global__ void __launch_bounds__(BAND_MAXW * 32, MINB)
normals_bwd_band(const float* __restrict__ d, const float* __restrict__ g, OP op, int H, int W, int rows, int chunks,
                 float* __restrict__ gd) {
    __shared__ float sx[2][BAND_MAXW][4];                             // [buffer][warp]: R of lane 31, L of lane 0, d.w of lane 31, d.x of lane 0
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    Task t;
    t.pl = blockIdx.x / chunks;
    t.i0 = (blockIdx.x - t.pl * chunks) * rows;
    t.i1 = min(t.i0 + rows, H);
    t.lane = lane;
    t.j = threadIdx.x * 4;
    t.act = t.j < W;
    const long plane = (long)H * W;
    const float* p = d + t.pl * plane;
    const float* gp = g + (long)t.pl * 3 * plane;
    float* o = gd + t.pl * plane;
    if (!op.begin(t.pl)) {                                            // block-uniform
        if constexpr (NEW) roll_generic_bwd(p, gp, op.cam, H, W, t, o, plane);
        return;
    }
    const RowPipe<4, 0, BAND_DEPTH, 0> pipe;
    const int jc = t.act ? t.j : 0;
    const int r0 = t.i0 - 1;
    const float* gp1 = gp + plane;
    const float* gp2 = gp1 + plane;
    auto issue = [&](int k, int slot) {                              // 32-bit offsets inside the plane (H * W < 2^31)
        if (k <= t.i1) {
            const int od = clampi(k + 1, H) * W + jc, og = clampi(k, H) * W + jc;
            const bool ga = (unsigned)k < (unsigned)H && t.act;
            pipe.cp16(slot, 0, p + od, t.act);
            pipe.cp16(slot, 1, gp + og, ga);
            pipe.cp16(slot, 2, gp1 + og, ga);
            pipe.cp16(slot, 3, gp2 + og, ga);
        }
        pipe.commit();
    };
#pragma unroll
    for (int k = 0; k < BAND_DEPTH; ++k) issue(r0 + k, k);
    float4 dU = ldrow(p, r0 - 1, H, W, t), dC = ldrow(p, r0, H, W, t);
    // column-border constants of the lane's two pairs (pixel 0 may be column 0, pixel 3 may be column W - 1)
    const bool first = t.j == 0, last = t.j + 4 >= W;                 // (W % 4 == 0)
    const f2 hmaA = mk(first ? 0.f : 0.5f, 0.5f), hmbA = bc(0.5f), hmabA = add(hmaA, hmbA);
    const f2 hmaB = bc(0.5f), hmbB = mk(0.5f, last ? 0.f : 0.5f), hmabB = add(hmaB, hmbB);
    const f2 fwA = mk(first ? 1.f : 0.5f, 0.5f), fwB = mk(0.5f, last ? 1.f : 0.5f);
    int buf = 0;
    if (lane == 31) sx[1][warp][2] = dC.w;
    if (lane == 0) sx[1][warp][3] = dC.x;
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, DnP[4] = {0.f, 0.f, 0.f, 0.f};   // acc = Hs(r - 1) + Dn(r - 2)
    op.row_begin(r0, t.j);
    int slot = 0;
    for (int r = r0; r <= t.i1; ++r) {
        pipe.wait();
        const float4 dD = pipe.v(slot, 0), G0 = pipe.v(slot, 1), G1 = pipe.v(slot, 2), G2 = pipe.v(slot, 3);
        issue(r + BAND_DEPTH, slot);
        slot = slot + 1 == BAND_DEPTH ? 0 : slot + 1;
        f2 RA = bc(0.f), LA = bc(0.f), DnA = bc(0.f), UpA = bc(0.f), RB = bc(0.f), LB = bc(0.f), DnB = bc(0.f), UpB = bc(0.f);
        float xl = __shfl_up_sync(FULL, dC.w, 1), xr = __shfl_down_sync(FULL, dC.x, 1);
        if (lane == 0) xl = warp > 0 ? sx[buf ^ 1][warp - 1][2] : dC.x;
        if (last) xr = dC.w; else if (lane == 31) xr = sx[buf ^ 1][warp + 1][3];
        if (r >= 0 && r < H) {                                        // block-uniform
            const float hmu_ = r > 0 ? 0.5f : 0.f, hmd_ = r < H - 1 ? 0.5f : 0.f, fh = edge_half(r, H);
            const f2 dlA = mk(xl, dC.x), drA = mk(dC.y, dC.z), dlB = drA, drB = mk(dC.w, xr);
            const f2 duA = mk(dU.x, dU.y), duB = mk(dU.z, dU.w), ddA = mk(dD.x, dD.y), ddB = mk(dD.z, dD.w);
            const f2 g0A = mk(G0.x, G0.y), g0B = mk(G0.z, G0.w), g1A = mk(G1.x, G1.y), g1B = mk(G1.z, G1.w), g2A = mk(G2.x, G2.y),
                     g2B = mk(G2.z, G2.w);
            if constexpr (NEW) {
                const float rx0 = (float)op.rxd, ry0 = (float)op.ryd;
                const f2 nrxA = mk(-rx0, -(rx0 + op.k0)), nryA = mk(-ry0, -(ry0 + op.k3));
                const f2 nrxB = mk(-(rx0 + 2.f * op.k0), -(rx0 + 3.f * op.k0)), nryB = mk(-(ry0 + 2.f * op.k3), -(ry0 + 3.f * op.k3));
                const f2 hmu = bc(hmu_), hmd = bc(hmd_), hmud = bc(hmu_ + hmd_), fhp = bc(fh);
                aff_adj_pair(op.k, nrxA, nryA, dlA, drA, duA, ddA, hmaA, hmbA, hmabA, hmu, hmd, hmud, mul(fhp, fwA), g0A, g1A, g2A, RA, LA,
                             DnA, UpA);
                aff_adj_pair(op.k, nrxB, nryB, dlB, drB, duB, ddB, hmaB, hmbB, hmabB, hmu, hmd, hmud, mul(fhp, fwB), g0B, g1B, g2B, RB, LB,
                             DnB, UpB);
            } else {
                old_adj_pair(op.scale, dlA, drA, duA, ddA, bc(fh), fwA, g0A, g1A, g2A, RA, LA, DnA, UpA);
                old_adj_pair(op.scale, dlB, drB, duB, ddB, bc(fh), fwB, g0B, g1B, g2B, RB, LB, DnB, UpB);
            }
        }
        // seam exchange: this row's R / L and the NEXT row's edge depths
        if (lane == 31) { sx[buf][warp][0] = RB.v.y; sx[buf][warp][2] = dD.w; }
        if (lane == 0) { sx[buf][warp][1] = LA.v.x; sx[buf][warp][3] = dD.x; }
        float Rl = __shfl_up_sync(FULL, RB.v.y, 1), Lr = __shfl_down_sync(FULL, LA.v.x, 1);
        __syncthreads();
        if (lane == 0) Rl = warp > 0 ? sx[buf][warp - 1][0] : 0.f;
        if (lane == 31) Lr = warp + 1 < nw ? sx[buf][warp + 1][1] : 0.f;
        buf ^= 1;
        float Hs[4] = {Rl + LA.v.y, RA.v.x + LB.v.x, RA.v.y + LB.v.y, RB.v.x + Lr};
        if (first) Hs[0] += LA.v.x;                                   // border-clamped neighbours fold onto the pixel itself
        if (last) Hs[3] += RB.v.y;
        if (r == 0) { Hs[0] += UpA.v.x; Hs[1] += UpA.v.y; Hs[2] += UpB.v.x; Hs[3] += UpB.v.y; }
        if (r == H - 1) { Hs[0] += DnA.v.x; Hs[1] += DnA.v.y; Hs[2] += DnB.v.x; Hs[3] += DnB.v.y; }
        if (r > t.i0 && t.act)                                        // row r - 1 is complete
            st4(o + ((r - 1) * W + t.j), make_float4(acc[0] + UpA.v.x, acc[1] + UpA.v.y, acc[2] + UpB.v.x, acc[3] + UpB.v.y));
        acc[0] = Hs[0] + DnP[0]; acc[1] = Hs[1] + DnP[1]; acc[2] = Hs[2] + DnP[2]; acc[3] = Hs[3] + DnP[3];
        DnP[0] = DnA.v.x; DnP[1] = DnA.v.y; DnP[2] = DnB.v.x; DnP[3] = DnB.v.y;
        dU = dC; dC = dD;
        op.row_next();
    }
}

// ------------------------------------------------------------------------------------------
// TV (main_model.py:15-19): sum of squared forward differences over (planes, H, W).  Pipe row k = row k + 1
// ------------------------------------------------------------------------------------------
#define TV_DEPTH 8
__global__ void __launch_bounds__(RNT, 8)
tv_fwd_roll(const float* __restrict__ x, int H, int W, RollPlan pl, double* __restrict__ out) {
    Task t;
    if (!roll_task(pl, H, W, t)) return;
    const float* p = x + (long)t.pl * H * W;
    const RowPipe<1, 1, TV_DEPTH> pipe;
    const bool seamR = t.lane == 31 && t.j + 4 < W;
    const int jc = t.act ? t.j : 0;
    auto issue = [&](int k, int slot) {
        if (k < t.i1) {
            const int od = clampi(k + 1, H) * W;
            pipe.cp16(slot, 0, p + (od + jc), t.act);
            if (seamR) pipe.cp4(slot, 0, p + (od + t.j + 4), true);
        }
        pipe.commit();
    };
#pragma unroll
    for (int k = 0; k < TV_DEPTH; ++k) issue(t.i0 + k, k);
    int slot = 0;
    float4 c = ldrow(p, t.i0, H, W, t);
    float sC = seamR ? __ldg(p + (long)t.i0 * W + t.j + 4) : 0.f;
    double acc = 0.0;
    float a = 0.f;
    for (int r = t.i0; r < t.i1; ++r) {
        pipe.wait();
        const float4 dn = pipe.v(slot, 0);
        const float sD = pipe.s(slot, 0);
        issue(r + TV_DEPTH, slot);
        slot = slot + 1 == TV_DEPTH ? 0 : slot + 1;
        float xr = __shfl_down_sync(FULL, c.x, 1);
        if (t.j + 4 >= W) xr = c.w; else if (t.lane == 31) xr = sC;
        const float h0 = c.x - c.y, h1 = c.y - c.z, h2 = c.z - c.w, h3 = c.w - xr;
        const float v0 = c.x - dn.x, v1 = c.y - dn.y, v2 = c.z - dn.z, v3 = c.w - dn.w;
        a += (h0 * h0 + v0 * v0) + (h1 * h1 + v1 * v1) + (h2 * h2 + v2 * v2) + (h3 * h3 + v3 * v3);
        if (((r - t.i0) & 15) == 15) { acc += (double)a; a = 0.f; }
        c = dn; sC = sD;
    }
    acc += (double)a;
    acc = warp_sum(acc);
    if (t.lane == 0) atomicAdd(out, acc);
}

// ------------------------------------------------------------------------------------------
// host side: called by the C-ABI entry points of stencil_tiled.cu; return 1 when the launch was made, 0 when the shape does
// not suit (ragged rows: the register-quad kernels take it), < 0 on error
// ------------------------------------------------------------------------------------------
static bool roll_ok(int W, const void* a, const void* b, const void* c) {
    return (W & 3) == 0 && W >= 4 && !(((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15);
}
#define ROLL_LAUNCH(kernel, smem, planes, ...)                                                      \
    do {                                                                                            \
        static int rw = 0;                                                                          \
        if (!rw) rw = resident_warps(kernel, smem);                                                 \
        RollPlan pl;                                                                                \
        if (!roll_plan(planes, H, W, rw, pl)) return 0;                                             \
        kernel<<<(pl.ntasks + RW - 1) / RW, RNT, smem, ST(stream)>>>(__VA_ARGS__);                  \
    } while (0)
// tuning knob (DSR_ROLL_MINB_<name> = 4 / 5 / 6 / 8 CTAs of 128 threads per SM, i.e. the register cap): measured defaults below
static int roll_minb(const char* name, int dflt) {
    char key[64];
    snprintf(key, sizeof(key), "DSR_ROLL_MINB_%s", name);
    const char* e = getenv(key);
    const int v = e ? atoi(e) : dflt;
    return (v == 2 || v == 3 || v == 4 || v == 5 || v == 6 || v == 8) ? v : dflt;
}
#define ROLL_DISPATCH(name, dflt, KT, smem, planes, ...)                                                \
    do {                                                                                            \
        static int mb = 0;                                                                          \
        if (!mb) mb = roll_minb(name, dflt);                                                        \
        if (mb == 4) { auto kern = KT(4); ROLL_LAUNCH(kern, smem, planes, __VA_ARGS__); }                 \
        else if (mb == 5) { auto kern = KT(5); ROLL_LAUNCH(kern, smem, planes, __VA_ARGS__); }            \
        else if (mb == 6) { auto kern = KT(6); ROLL_LAUNCH(kern, smem, planes, __VA_ARGS__); }            \
        else { auto kern = KT(8); ROLL_LAUNCH(kern, smem, planes, __VA_ARGS__); }                         \
    } while (0)


// row-band launches: grid = planes x chunks CTAs of ceil(W / 128) warps, chunks sized for whole waves of resident CTAs
template <typename K>
static int band_launch_plan(K kernel, int nt, int smem, long planes, int H, int& rows, int& chunks) {
    int nb = 0;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RowPipe<4, 0, BAND_DEPTH, 0>::bytes(BAND_MAXW * 32));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, nt, smem) != cudaSuccess || nb < 1) nb = 1;
    const long slots = (long)dsr_num_sms() * nb;
    long ch = 1;
    for (int k = 1; k <= 64; ++k) { ch = k * slots / planes; if (ch >= 1) break; }
    if (ch < 1) ch = 1;
    long rw = (H + ch - 1) / ch;
    if (rw < 8) rw = 8;
    if (rw > H) rw = H;
    rows = (int)rw;
    chunks = (int)((H + rw - 1) / rw);
    return planes * chunks < (1L << 31) ? 1 : 0;
}
#define BAND_DISPATCH(name, dflt, KT, planes, ...)                                                  \
    do {                                                                                            \
        static int mb = 0;                                                                          \
        if (!mb) { mb = roll_minb(name, dflt); if (mb > 4) mb = 4; }                                \
        const int nt = ((W + RCOLS - 1) / RCOLS) * 32, smem = RowPipe<4, 0, BAND_DEPTH, 0>::bytes(nt); \
        int rows, chunks;                                                                           \
        if (mb == 2) { auto kern = KT(2); if (!band_launch_plan(kern, nt, smem, planes, H, rows, chunks)) return 0;      \
            kern<<<(unsigned)(planes * chunks), nt, smem, ST(stream)>>>(__VA_ARGS__); }             \
        else if (mb == 3) { auto kern = KT(3); if (!band_launch_plan(kern, nt, smem, planes, H, rows, chunks)) return 0; \
            kern<<<(unsigned)(planes * chunks), nt, smem, ST(stream)>>>(__VA_ARGS__); }             \
        else { auto kern = KT(4); if (!band_launch_plan(kern, nt, smem, planes, H, rows, chunks)) return 0;              \
            kern<<<(unsigned)(planes * chunks), nt, smem, ST(stream)>>>(__VA_ARGS__); }             \
    } while (0)

int dsr_roll_normals_old_fwd(const float* d, int B, int H, int W, float scale, float* out, void* stream) {
    if (!roll_ok(W, d, out, nullptr)) return 0;
    OldOp op; op.scale = scale;
#define KT_OF(m) normals_fwd_roll<OldOp, false, m>
    ROLL_DISPATCH("OLD_FWD", 8, KT_OF, (RowPipe<1, 1, FWD_DEPTH>::bytes()), B, d, op, H, W, pl, out);
#undef KT_OF
    return 1;
}
int dsr_roll_normals_old_bwd(const float* d, const float* g, int B, int H, int W, float scale, float* gd, void* stream) {
    if (!roll_ok(W, d, g, gd) || W > RCOLS * BAND_MAXW) return 0;
    OldBand op; op.scale = scale; op.cam = nullptr;
#define KT_OB(m) normals_bwd_band<OldBand, false, m>
    BAND_DISPATCH("OLD_BWD", 3, KT_OB, (long)B, d, g, op, H, W, rows, chunks, gd);
#undef KT_OB
    return 1;
}
int dsr_roll_normals_new_fwd(const float* d, const double* cams, int B, int H, int W, float* out, void* stream) {
    if (!roll_ok(W, d, out, nullptr)) return 0;
    NewOp op; op.cams = cams;
#define KT_NF(m) normals_fwd_roll<NewOp, true, m>
    ROLL_DISPATCH("NEW_FWD", 6, KT_NF, (RowPipe<1, 1, FWD_DEPTH>::bytes()), B, d, op, H, W, pl, out);
#undef KT_NF
    return 1;
}
int dsr_roll_normals_new_bwd(const float* d, const float* g, const double* cams, int B, int H, int W, float* gd, void* stream) {
    if (!roll_ok(W, d, g, gd) || W > RCOLS * BAND_MAXW) return 0;
    NewBand op; op.cams = cams; op.cam = nullptr;
#define KT_NB(m) normals_bwd_band<NewBand, true, m>
    BAND_DISPATCH("NEW_BWD", 3, KT_NB, (long)B, d, g, op, H, W, rows, chunks, gd);
#undef KT_NB
    return 1;
}
int dsr_roll_tv_fwd(const float* x, long planes, int H, int W, double* out, void* stream) {
    if (!roll_ok(W, x, nullptr, nullptr)) return 0;
    ROLL_LAUNCH(tv_fwd_roll, (RowPipe<1, 1, TV_DEPTH>::bytes()), planes, x, H, W, pl, out);
    return 1;
}
