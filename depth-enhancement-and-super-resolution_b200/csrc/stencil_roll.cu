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
//     R / L travel between lanes by shuffle, Dn / Up between rows in registers; the backward kernels run as ROW BANDS (one
//     CTA = all strips of a band of rows) so that the seam values cross warps through 16 floats of shared memory and one
//     barrier per row (the first strip version recomputed the seam pixels in lanes 0 / 31: 100 of 660 instructions per row);
//   * loads travel through per-thread cp.async slots (RowPipe), math runs on packed fp32 pairs (FFMA2 / FMUL2);
//   * the B * H rows are cut into chunks so that all tasks fit whole waves of resident warps / CTAs (no tail wave).
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

// per-pixel fp64 reference form for cameras with a perspective row in K^-1 (never a pin-hole K): slow, exact, block-uniform
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
    uint32_t vbase, sbase, ustride, wstride;                          // byte strides between units / words of one slot
    static constexpr int bytes(int nt = NT) { return DEPTH * (NV * 16 + NS * 4) * nt; }
    __device__ __forceinline__ RowPipe() {
        extern __shared__ __align__(16) unsigned char roll_smem[];
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(roll_smem);
        const int nt = NT ? NT : (int)blockDim.x;                     // threads per CTA (NT == 0: blockDim.x)
        ustride = nt * 16; wstride = nt * 4;
        vbase = b + threadIdx.x * 16;
        sbase = b + DEPTH * NV * 16 * nt + threadIdx.x * 4;
    }
    __device__ __forceinline__ uint32_t vslot(int slot) const { return vbase + slot * NV * ustride; }
    __device__ __forceinline__ uint32_t sslot(int slot) const { return sbase + slot * NS * wstride; }
    __device__ __forceinline__ void cp16(int slot, int u, const float* src, bool valid) const {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vslot(slot) + u * ustride), "l"(src), "r"(valid ? 16 : 0));
    }
    __device__ __forceinline__ void cp4(int slot, int w, const float* src, bool valid) const {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sslot(slot) + w * wstride), "l"(src), "r"(valid ? 4 : 0));
    }
    __device__ __forceinline__ float4 v(int slot, int u) const {
        float4 r;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(vslot(slot) + u * ustride));
        return r;
    }
    __device__ __forceinline__ float s(int slot, int w) const {
        float r;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(sslot(slot) + w * wstride));
        return r;
    }
    __device__ __forceinline__ static void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    __device__ __forceinline__ static void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory"); }
};
__device__ __forceinline__ int clampi(int v, int n) { return min(max(v, 0), n - 1); }

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

// per-plane constants of the camera-space normals, as broadcast pairs
struct AffPair {
    f2 k0, k4, nk1, nk3, D;
    __device__ __forceinline__ void set(const AffCam& c) { k0 = bc(c.k0); k4 = bc(c.k4); nk1 = bc(-c.k1); nk3 = bc(-c.k3); D = bc(c.D); }
};
// The closed form of stencil_math.cuh (aff_terms / aff_normal_m) with its POWER-OF-TWO factors dropped: m is bilinear in
// (Du, Su, Dv, Sv), each of which carries np.gradient's / the depth map's 1/2, and `sc` (1, 1/2 or 1/4) multiplies all of m,
// so m here = 4 m_ref / sc EXACTLY (scaling by a power of two), the unit normal is unchanged, and in the adjoint the factors
// cancel between dn/dm (1 / |m|) and dm/dd.  Only the clamp of norms.py:103 (|m_ref| < 1e-12) sees the scale: its threshold
// becomes thr = 16e-24 / sc^2 on |m|^2, which reproduces n = m_ref * 1e12 for degenerate pixels.
// Borders without branches: ma / mb = 1 when the left / right neighbour is a real pixel (0: clamped onto the pixel itself),
// mab = ma + mb, likewise mu / md / mud for the rows.
__device__ __forceinline__ void aff_m_pair(const AffPair& k, f2 nrx, f2 nry, f2 dl, f2 dr, f2 du, f2 dd, f2 ma, f2 mb, f2 mab, f2 mu,
                                           f2 md, f2 mud, f2& Du, f2& Su, f2& Dv, f2& Sv, f2& m0, f2& m1, f2& m2) {
    Du = sub(dr, dl);
    Su = fma(mb, dr, fma(ma, dl, mab));
    Dv = sub(dd, du);
    Sv = fma(md, dd, fma(mu, du, mud));
    const f2 A = mul(Su, Dv), B = mul(Sv, Du);
    m0 = fma(k.k4, B, mul(k.nk3, A));
    m1 = fma(k.k0, A, mul(k.nk1, B));
    m2 = fma(nrx, m0, fma(nry, m1, mul(k.D, mul(Su, Sv))));
}
__device__ __forceinline__ f2 inv_len_thr(f2 r2, f2 thr) { return mk(rsqrtf(fmaxf(r2.v.x, thr.v.x)), rsqrtf(fmaxf(r2.v.y, thr.v.y))); }
__device__ __forceinline__ void aff_fwd_pair(const AffPair& k, f2 nrx, f2 nry, f2 dl, f2 dr, f2 du, f2 dd, f2 ma, f2 mb, f2 mab, f2 mu,
                                             f2 md, f2 mud, f2 thr, f2& n0, f2& n1, f2& n2) {
    f2 Du, Su, Dv, Sv, m0, m1, m2;
    aff_m_pair(k, nrx, nry, dl, dr, du, dd, ma, mb, mab, mu, md, mud, Du, Su, Dv, Sv, m0, m1, m2);
    const f2 ir = inv_len_thr(fma(m0, m0, fma(m1, m1, mul(m2, m2))), thr);
    n0 = mul(m0, ir); n1 = mul(m1, ir); n2 = mul(m2, ir);
}
// adjoint of a pair: dL/dn (g) -> what each pixel adds to dL/dd of its right / left / lower / upper neighbour
__device__ __forceinline__ void aff_adj_pair(const AffPair& k, f2 nrx, f2 nry, f2 dl, f2 dr, f2 du, f2 dd, f2 ma, f2 mb, f2 mab, f2 mu,
                                             f2 md, f2 mud, f2 thr, f2 g0, f2 g1, f2 g2, f2& R, f2& L, f2& Dn, f2& Up) {
    f2 Du, Su, Dv, Sv, m0, m1, m2;
    aff_m_pair(k, nrx, nry, dl, dr, du, dd, ma, mb, mab, mu, md, mud, Du, Su, Dv, Sv, m0, m1, m2);
    const f2 r2 = fma(m0, m0, fma(m1, m1, mul(m2, m2)));
    const f2 ir = inv_len_thr(r2, thr);
    const f2 n0 = mul(m0, ir), n1 = mul(m1, ir), n2 = mul(m2, ir);
    const f2 dot = fma(n0, g0, fma(n1, g1, mul(n2, g2)));
    const f2 ndot = mk(r2.v.x > thr.v.x ? -dot.v.x : 0.f, r2.v.y > thr.v.y ? -dot.v.y : 0.f);   // clamped pixels: no projection
    const f2 dm0 = mul(fma(n0, ndot, g0), ir), dm1 = mul(fma(n1, ndot, g1), ir), dm2 = mul(fma(n2, ndot, g2), ir);
    const f2 e0 = fma(nrx, dm2, dm0), e1 = fma(nry, dm2, dm1), e2 = mul(k.D, dm2);
    const f2 p = fma(k.k4, e0, mul(k.nk1, e1)), q = fma(k.k0, e1, mul(k.nk3, e0));
    const f2 dDu = mul(Sv, p), dSv = fma(Du, p, mul(e2, Su)), dDv = mul(Su, q), dSu = fma(Dv, q, mul(e2, Sv));
    R = fma(mb, dSu, dDu);
    L = sub(mul(ma, dSu), dDu);
    Dn = fma(md, dSv, dDv);
    Up = sub(mul(mu, dSv), dDv);
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
// forward, camera-space normals: depth (planes, H, W) -> normals (planes, 3, H, W).  Warp strips; pipe row k = depth row k + 1
// (+ the seam lane's neighbour column).  (The image-space forward kernel stays on the register quads of stencil_tiled.cu:
// 76 % of the HBM peak there, and its IEEE sqrt / division - kept for the 1e-4 gate on values of O(100) - make the strip form
// no faster: measured r2c 61 %.)
// ------------------------------------------------------------------------------------------
#define FWD_DEPTH 6
template <int MINB>
__global__ void __launch_bounds__(RNT, MINB)
normals_new_fwd_roll(const float* __restrict__ d, NewBand op, int H, int W, RollPlan pl, float* __restrict__ out) {
    Task t;
    if (!roll_task(pl, H, W, t)) return;
    const long plane = (long)H * W;
    const float* p = d + t.pl * plane;
    float* o = out + (long)t.pl * 3 * plane;
    if (!op.begin(t.pl)) {
        roll_generic_fwd(p, op.cam, H, W, t, o, plane);
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
    const bool first = t.j == 0, last = t.j + 4 >= W;                 // (W % 4 == 0)
    const f2 maA = mk(first ? 0.f : 1.f, 1.f), mbA = bc(1.f), mabA = add(maA, mbA);
    const f2 maB = bc(1.f), mbB = mk(1.f, last ? 0.f : 1.f), mabB = add(maB, mbB);
    const f2 ifw2A = mk(first ? 16e-24f : 64e-24f, 64e-24f), ifw2B = mk(64e-24f, last ? 16e-24f : 64e-24f);   // 16e-24 / fw^2
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
        if (last) xr = dC.w; else if (t.lane == 31) xr = sC;
        const float mu_ = r > 0 ? 1.f : 0.f, md_ = r < H - 1 ? 1.f : 0.f;
        const f2 mu = bc(mu_), md = bc(md_), mud = bc(mu_ + md_), ifh2 = bc((r == 0 || r == H - 1) ? 1.f : 4.f);   // 1 / fh^2
        const f2 dlA = mk(xl, dC.x), drA = mk(dC.y, dC.z), dlB = drA, drB = mk(dC.w, xr);
        const f2 duA = mk(dU.x, dU.y), duB = mk(dU.z, dU.w), ddA = mk(dD.x, dD.y), ddB = mk(dD.z, dD.w);
        const float rx0 = (float)op.rxd, ry0 = (float)op.ryd;
        const f2 nrxA = mk(-rx0, -(rx0 + op.k0)), nryA = mk(-ry0, -(ry0 + op.k3));
        const f2 nrxB = mk(-(rx0 + 2.f * op.k0), -(rx0 + 3.f * op.k0)), nryB = mk(-(ry0 + 2.f * op.k3), -(ry0 + 3.f * op.k3));
        f2 n0A, n1A, n2A, n0B, n1B, n2B;
        aff_fwd_pair(op.k, nrxA, nryA, dlA, drA, duA, ddA, maA, mbA, mabA, mu, md, mud, mul(ifh2, ifw2A), n0A, n1A, n2A);
        aff_fwd_pair(op.k, nrxB, nryB, dlB, drB, duB, ddB, maB, mbB, mabB, mu, md, mud, mul(ifh2, ifw2B), n0B, n1B, n2B);
        if (t.act) {
            float* q = o + (r * W + t.j);
            st4(q, make_float4(n0A.v.x, n0A.v.y, n0B.v.x, n0B.v.y));
            st4(q + plane, make_float4(n1A.v.x, n1A.v.y, n1B.v.x, n1B.v.y));
            st4(q + 2 * plane, make_float4(n2A.v.x, n2A.v.y, n2B.v.x, n2B.v.y));
        }
        dU = dC; dC = dD; sC = sD;
        op.row_next();
    }
}

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
    const f2 maA = mk(first ? 0.f : 1.f, 1.f), mbA = bc(1.f), mabA = add(maA, mbA);
    const f2 maB = bc(1.f), mbB = mk(1.f, last ? 0.f : 1.f), mabB = add(maB, mbB);
    const f2 fwA = mk(first ? 1.f : 0.5f, 0.5f), fwB = mk(0.5f, last ? 1.f : 0.5f);
    const f2 ifw2A = mk(first ? 16e-24f : 64e-24f, 64e-24f), ifw2B = mk(64e-24f, last ? 16e-24f : 64e-24f);   // 16e-24 / fw^2
    int buf = 0;
    if (lane == 31) sx[1][warp][2] = dC.w;
    if (lane == 0) sx[1][warp][3] = dC.x;
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, DnP[4] = {0.f, 0.f, 0.f, 0.f};   // acc = Hs(r - 1) + Dn(r - 2)
    float* orow = o + (t.i0 * W + jc);
    // negated rays of the lane's two pairs: evaluated in fp64 for the chunk's first row, then advanced by k1 / k4 per row in
    // fp32 (<= 2^-24 |ray| per step over the <= ~100 rows of a chunk: 1e-6 relative on a GRADIENT whose gate is 1e-3; the
    // forward kernel, gated at 2e-6 absolute, keeps the fp64 advance)
    f2 nrxA = bc(0.f), nryA = bc(0.f), nrxB = bc(0.f), nryB = bc(0.f), nk1p = bc(0.f), nk4p = bc(0.f);
    if constexpr (NEW) {
        op.row_begin(r0, t.j);
        const float rx0 = (float)op.rxd, ry0 = (float)op.ryd;
        nrxA = mk(-rx0, -(rx0 + op.k0)); nryA = mk(-ry0, -(ry0 + op.k3));
        nrxB = mk(-(rx0 + 2.f * op.k0), -(rx0 + 3.f * op.k0)); nryB = mk(-(ry0 + 2.f * op.k3), -(ry0 + 3.f * op.k3));
        nk1p = bc(-(float)op.k1d); nk4p = bc(-(float)op.k4d);
    }
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
            const float mu_ = r > 0 ? 1.f : 0.f, md_ = r < H - 1 ? 1.f : 0.f, fh = edge_half(r, H);
            const f2 dlA = mk(xl, dC.x), drA = mk(dC.y, dC.z), dlB = drA, drB = mk(dC.w, xr);
            const f2 duA = mk(dU.x, dU.y), duB = mk(dU.z, dU.w), ddA = mk(dD.x, dD.y), ddB = mk(dD.z, dD.w);
            const f2 g0A = mk(G0.x, G0.y), g0B = mk(G0.z, G0.w), g1A = mk(G1.x, G1.y), g1B = mk(G1.z, G1.w), g2A = mk(G2.x, G2.y),
                     g2B = mk(G2.z, G2.w);
            if constexpr (NEW) {
                const f2 mu = bc(mu_), md = bc(md_), mud = bc(mu_ + md_), ifh2 = bc(fh == 1.f ? 1.f : 4.f);       // 1 / fh^2
                aff_adj_pair(op.k, nrxA, nryA, dlA, drA, duA, ddA, maA, mbA, mabA, mu, md, mud, mul(ifh2, ifw2A), g0A, g1A, g2A, RA, LA,
                             DnA, UpA);
                aff_adj_pair(op.k, nrxB, nryB, dlB, drB, duB, ddB, maB, mbB, mabB, mu, md, mud, mul(ifh2, ifw2B), g0B, g1B, g2B, RB, LB,
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
        if (r > t.i0) {                                               // row r - 1 is complete
            if (t.act) st4(orow, make_float4(acc[0] + UpA.v.x, acc[1] + UpA.v.y, acc[2] + UpB.v.x, acc[3] + UpB.v.y));
            orow += W;
        }
        acc[0] = Hs[0] + DnP[0]; acc[1] = Hs[1] + DnP[1]; acc[2] = Hs[2] + DnP[2]; acc[3] = Hs[3] + DnP[3];
        DnP[0] = DnA.v.x; DnP[1] = DnA.v.y; DnP[2] = DnB.v.x; DnP[3] = DnB.v.y;
        dU = dC; dC = dD;
        if constexpr (NEW) { nrxA = add(nrxA, nk1p); nrxB = add(nrxB, nk1p); nryA = add(nryA, nk4p); nryB = add(nryB, nk4p); }
    }
}

// ------------------------------------------------------------------------------------------
// edge-aware smoothness, one pyramid level (main_model.py:22-73): d (B, 1, h, w), img (B, C, h, w).  Per EDGE between two
// pixels: w = exp(-mean_c |I_a - I_b|); forward sum += |(d_a - d_b) w|; backward t = coef * w * sign((d_a - d_b) w),
// gd_a += t, gd_b -= t.  One thread = 4 columns, rolling down the rows: it owns the 4 vertical edges below and the 4
// horizontal edges right of its pixels; the edge left of its first pixel comes from the neighbouring lane (shuffle; the seam
// lane recomputes that one edge from its neighbour column, which travels through the pipe like the rows do).
// Pipe row k = row k + 1 of the depth plane and of the C image planes (+ the seam lane's neighbour column of each).
// ------------------------------------------------------------------------------------------
#define SM_DEPTH 3
__device__ __forceinline__ float signed_or_zero(float mag, float p) { return p == 0.f ? 0.f : copysignf(mag, p); }
template <int C, bool BWD>
__global__ void __launch_bounds__(RNT, 6)
smooth_roll(const float* __restrict__ d, const float* __restrict__ img, int H, int W, RollPlan pl, const float* __restrict__ gscale,
            float cx, float cy, float* __restrict__ gd, double* __restrict__ out) {
    Task t;
    if (!roll_task(pl, H, W, t)) return;
    const long plane = (long)H * W;
    const float* p[C + 1];
    p[0] = d + t.pl * plane;
#pragma unroll
    for (int c = 0; c < C; ++c) p[c + 1] = img + ((long)t.pl * C + c) * plane;
    const RowPipe<C + 1, C + 1, SM_DEPTH> pipe;
    const bool seamL = BWD && t.lane == 0 && t.j > 0, seamR = t.lane == 31 && t.j + 4 < W, seam = seamL || seamR;
    const int jn = seamL ? t.j - 1 : (seamR ? t.j + 4 : 0);
    const int jc = t.act ? t.j : 0;
    const bool last = t.j + 4 >= W;
    const int r0 = BWD ? max(t.i0 - 1, 0) : t.i0;                     // backward: the vertical edges above the chunk's first row
    auto issue = [&](int k, int slot) {
        if (k < t.i1) {
            const int od = clampi(k + 1, H) * W;
#pragma unroll
            for (int c = 0; c <= C; ++c) {
                pipe.cp16(slot, c, p[c] + (od + jc), t.act);
                if (seam) pipe.cp4(slot, c, p[c] + (od + jn), true);
            }
        }
        pipe.commit();
    };
#pragma unroll
    for (int k = 0; k < SM_DEPTH; ++k) issue(r0 + k, k);
    float4 cur[C + 1];
    float sC[C + 1];
#pragma unroll
    for (int c = 0; c <= C; ++c) {
        cur[c] = ldrow(p[c], r0, H, W, t);
        sC[c] = seam ? __ldg(p[c] + (long)r0 * W + jn) : 0.f;
    }
    const float invC = 1.f / (float)C, g = BWD ? (gscale ? *gscale : 1.f) : 0.f, gcx = g * cx, gcy = g * cy;
    float tvp[4] = {0.f, 0.f, 0.f, 0.f};                              // backward: t of the vertical edges above the current row
    float fx = 0.f, fy = 0.f;
    double ax = 0.0, ay = 0.0;
    float* orow = BWD ? gd + t.pl * plane + (t.i0 * W + jc) : nullptr;
    int slot = 0;
    for (int r = r0; r < t.i1; ++r) {
        pipe.wait();
        float4 nxt[C + 1];
        float sD[C + 1];
#pragma unroll
        for (int c = 0; c <= C; ++c) { nxt[c] = pipe.v(slot, c); sD[c] = pipe.s(slot, c); }
        issue(r + SM_DEPTH, slot);
        slot = slot + 1 == SM_DEPTH ? 0 : slot + 1;
        float sv[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f}, shl = 0.f, dright, dleft = 0.f;
#pragma unroll
        for (int c = 0; c <= C; ++c) {
            const float4 a = cur[c], b = nxt[c];
            float xr = __shfl_down_sync(FULL, a.x, 1);
            if (last) xr = a.w; else if (t.lane == 31) xr = sC[c];          // clamped: the edge across the border is 0
            if (c == 0) {
                dright = xr;
                if (seamL) dleft = sC[0];
            } else {
                sv[0] += fabsf(a.x - b.x); sv[1] += fabsf(a.y - b.y); sv[2] += fabsf(a.z - b.z); sv[3] += fabsf(a.w - b.w);
                sh[0] += fabsf(a.x - a.y); sh[1] += fabsf(a.y - a.z); sh[2] += fabsf(a.z - a.w); sh[3] += fabsf(a.w - xr);
                if (seamL) shl += fabsf(sC[c] - a.x);
            }
        }
        const float4 dc = cur[0], dn = nxt[0];
        const float dv[4] = {dc.x - dn.x, dc.y - dn.y, dc.z - dn.z, dc.w - dn.w};
        const float dh[4] = {dc.x - dc.y, dc.y - dc.z, dc.z - dc.w, dc.w - dright};
        if (!BWD) {
            if (t.act) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    fx += fabsf(dv[e] * __expf(-sv[e] * invC));
                    fy += fabsf(dh[e] * __expf(-sh[e] * invC));
                }
            }
            if (((r - r0) & 15) == 15) { ax += (double)fx; ay += (double)fy; fx = 0.f; fy = 0.f; }
        } else {
            float tv[4], th[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float wv = __expf(-sv[e] * invC), wh = __expf(-sh[e] * invC);
                tv[e] = signed_or_zero(gcx * wv, dv[e] * wv);
                th[e] = signed_or_zero(gcy * wh, dh[e] * wh);
            }
            float tl = __shfl_up_sync(FULL, th[3], 1);                // the edge left of the lane's first pixel
            if (t.lane == 0) {
                tl = 0.f;
                if (seamL) { const float wl = __expf(-shl * invC); tl = signed_or_zero(gcy * wl, (dleft - dc.x) * wl); }
            }
            if (r >= t.i0) {
                if (t.act) st4(orow, make_float4(tv[0] - tvp[0] + th[0] - tl, tv[1] - tvp[1] + th[1] - th[0], tv[2] - tvp[2] + th[2] - th[1],
                                                 tv[3] - tvp[3] + th[3] - th[2]));
                orow += W;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) tvp[e] = tv[e];
        }
#pragma unroll
        for (int c = 0; c <= C; ++c) { cur[c] = nxt[c]; sC[c] = sD[c]; }
    }
    if (!BWD) {
        ax += (double)fx; ay += (double)fy;
        ax = warp_sum(ax); ay = warp_sum(ay);
        if (t.lane == 0) { atomicAdd(out, ax); atomicAdd(out + 1, ay); }
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

int dsr_roll_normals_old_fwd(const float*, int, int, int, float, float*, void*) { return 0; }   // stays on the register quads
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
    NewBand op; op.cams = cams; op.cam = nullptr;
#define KT_NF(m) normals_new_fwd_roll<m>
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
template <int C, bool BWD>
static int smooth_roll_launch(const float* d, const float* img, int B, int H, int W, const float* gscale, float cx, float cy, float* gd,
                              double* out, void* stream) {
    auto kern = smooth_roll<C, BWD>;
    ROLL_LAUNCH(kern, (RowPipe<C + 1, C + 1, SM_DEPTH>::bytes()), (long)B, d, img, H, W, pl, gscale, cx, cy, gd, out);
    return 1;
}
int dsr_roll_smooth_fwd(const float* d, const float* img, int B, int C, int H, int W, double* out2, void* stream) {
    if (!roll_ok(W, d, img, nullptr)) return 0;
    switch (C) {
        case 1: return smooth_roll_launch<1, false>(d, img, B, H, W, nullptr, 0.f, 0.f, nullptr, out2, stream);
        case 2: return smooth_roll_launch<2, false>(d, img, B, H, W, nullptr, 0.f, 0.f, nullptr, out2, stream);
        case 3: return smooth_roll_launch<3, false>(d, img, B, H, W, nullptr, 0.f, 0.f, nullptr, out2, stream);
        case 4: return smooth_roll_launch<4, false>(d, img, B, H, W, nullptr, 0.f, 0.f, nullptr, out2, stream);
    }
    return 0;
}
int dsr_roll_smooth_bwd(const float* d, const float* img, int B, int C, int H, int W, const float* gscale, float cx, float cy, float* gd,
                        void* stream) {
    if (!roll_ok(W, d, img, gd)) return 0;
    switch (C) {
        case 1: return smooth_roll_launch<1, true>(d, img, B, H, W, gscale, cx, cy, gd, nullptr, stream);
        case 2: return smooth_roll_launch<2, true>(d, img, B, H, W, gscale, cx, cy, gd, nullptr, stream);
        case 3: return smooth_roll_launch<3, true>(d, img, B, H, W, gscale, cx, cy, gd, nullptr, stream);
        case 4: return smooth_roll_launch<4, true>(d, img, B, H, W, gscale, cx, cy, gd, nullptr, stream);
    }
    return 0;
}
int dsr_roll_tv_fwd(const float* x, long planes, int H, int W, double* out, void* stream) {
    if (!roll_ok(W, x, nullptr, nullptr)) return 0;
    ROLL_LAUNCH(tv_fwd_roll, (RowPipe<1, 1, TV_DEPTH>::bytes()), planes, x, H, W, pl, out);
    return 1;
}
