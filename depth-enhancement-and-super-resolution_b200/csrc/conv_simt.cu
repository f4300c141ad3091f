// Generic fp32 implicit-GEMM convolution on CUDA cores (NHWC).  This is the always-correct path for
// every layer shape of models/networks.py / models/translation_network.py (any R,S, stride, zero
// padding, transposed or not, any channel count); the tcgen05 kernels in conv_tc.cu take over the
// tensor-core-shaped layers.  One gather formulation serves all six conv/convT fwd/dgrad/wgrad cases:
//
//   rows  m = (n, oh, ow) of a "base grid" (Ho x Wo);  k = (r, s, cg) over the "gathered" tensor G
//   normal     : ih = oh*stride - pad + r                       (Conv2d fwd, ConvTranspose2d dgrad)
//   transposed : ih = (oh + pad - r) / stride when divisible    (ConvTranspose2d fwd, Conv2d dgrad)
//   out[m][co] = bias[co] + sum_k G(m,k) * Wk[k][co]            (conv_simt_kernel)
//   dWk[k][cd] = sum_m G(m,k) * D[m][cd]                        (wgrad_simt_kernel, split over m)
#include "common.cuh"
#include "../../include/dsr_b200.h"

#define ST(s) ((cudaStream_t)(s))
#define BM 64
#define BN 64
#define BK 16

struct Geom {
    int N, Hg, Wg, Cg;  // gathered tensor
    int Ho, Wo;         // base grid
    int R, S, stride, pad, transposed;
};

__device__ __forceinline__ bool gather_coord(const Geom& g, int oh, int ow, int r, int s, int& ih, int& iw) {
    if (!g.transposed) {
        ih = oh * g.stride - g.pad + r;
        iw = ow * g.stride - g.pad + s;
        return ih >= 0 && ih < g.Hg && iw >= 0 && iw < g.Wg;
    }
    int th = oh + g.pad - r, tw = ow + g.pad - s;
    if (th < 0 || tw < 0) return false;
    if (g.stride > 1 && ((th % g.stride) || (tw % g.stride))) return false;
    ih = th / g.stride;
    iw = tw / g.stride;
    return ih < g.Hg && iw < g.Wg;
}

// VEC: Cg % 16 == 0, so a BK-wide k tile lies inside one filter tap and float4 loads are aligned
template <bool VEC>
__global__ void __launch_bounds__(256) conv_simt_kernel(const float* __restrict__ G, const float* __restrict__ Wk,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         Geom g, int Co, int act_out) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const long M = (long)g.N * g.Ho * g.Wo;
    const int K = g.R * g.S * g.Cg;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int arow = tid >> 2, akq = (tid & 3) * 4;
    const int brow = tid >> 4, bcol = (tid & 15) * 4;
    const int ty = tid >> 4, tx = tid & 15;
    long am = m0 + arow;
    int an = 0, aoh = 0, aow = 0;
    if (am < M) { aow = (int)(am % g.Wo); long t = am / g.Wo; aoh = (int)(t % g.Ho); an = (int)(t / g.Ho); }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        if (am < M) {
            if (VEC) {
                int k = k0 + akq;
                int tap = k / g.Cg, cg = k - tap * g.Cg;
                int r = tap / g.S, s = tap - r * g.S, ih, iw;
                if (gather_coord(g, aoh, aow, r, s, ih, iw)) {
                    float4 v = ld4(G + (((long)an * g.Hg + ih) * g.Wg + iw) * g.Cg + cg);
                    av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int k = k0 + akq + e;
                    if (k < K) {
                        int tap = k / g.Cg, cg = k - tap * g.Cg;
                        int r = tap / g.S, s = tap - r * g.S, ih, iw;
                        if (gather_coord(g, aoh, aow, r, s, ih, iw))
                            av[e] = G[(((long)an * g.Hg + ih) * g.Wg + iw) * g.Cg + cg];
                    }
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) As[akq + e][arow] = av[e];
        {
            int k = k0 + brow;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int co = n0 + bcol + e;
                Bs[brow][bcol + e] = (k < K && co < Co) ? Wk[(long)k * Co + co] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = n0 + tx * 4 + j;
            if (co >= Co) continue;
            float v = acc[i][j] + (bias ? bias[co] : 0.f);
            if (act_out == DSR_ACT_TANH) v = tanhf(v);
            out[m * Co + co] = v;
        }
    }
}

// dWk[k][cd] += sum over this CTA's m-range of G(m,k) * D[m][cd]
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const float* __restrict__ G, const float* __restrict__ D,
                                                          float* __restrict__ dWk, Geom g, int Cd, long m_per_split) {
    __shared__ __align__(16) float As[BK][BM + 4];   // [mm][k]
    __shared__ __align__(16) float Ds[BK][BN];       // [mm][cd]
    const int tid = threadIdx.x;
    const long M = (long)g.N * g.Ho * g.Wo;
    const int K = g.R * g.S * g.Cg;
    const int k0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    long m_begin = (long)blockIdx.z * m_per_split, m_end = m_begin + m_per_split;
    if (m_end > M) m_end = M;
    const int lrow = tid >> 4, lq = (tid & 15) * 4;
    const int ty = tid >> 4, tx = tid & 15;
    // the 4 k's this thread gathers are fixed for the whole kernel: decode their taps once
    int tr[4], ts[4], tc[4]; bool tv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int k = k0 + lq + e;
        tv[e] = k < K;
        int tap = tv[e] ? k / g.Cg : 0;
        tc[e] = tv[e] ? k - tap * g.Cg : 0;
        tr[e] = tap / g.S; ts[e] = tap - tr[e] * g.S;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (long mb = m_begin; mb < m_end; mb += BK) {
        long m = mb + lrow;
        float av[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < m_end) {
            int ow = (int)(m % g.Wo); long t = m / g.Wo; int oh = (int)(t % g.Ho); int n = (int)(t / g.Ho);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int ih, iw;
                if (tv[e] && gather_coord(g, oh, ow, tr[e], ts[e], ih, iw))
                    av[e] = G[(((long)n * g.Hg + ih) * g.Wg + iw) * g.Cg + tc[e]];
                int cd = n0 + lq + e;
                if (cd < Cd) dv[e] = D[m * Cd + cd];
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) { As[lrow][lq + e] = av[e]; Ds[lrow][lq + e] = dv[e]; }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < BK; ++mm) {
            float4 a4 = *reinterpret_cast<const float4*>(&As[mm][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&Ds[mm][tx * 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int k = k0 + ty * 4 + i;
        if (k >= K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int cd = n0 + tx * 4 + j;
            if (cd < Cd) atomicAdd(&dWk[(long)k * Cd + cd], acc[i][j]);
        }
    }
}

static int fill_geom(Geom* g, int N, int Hg, int Wg, int Cg, int Ho, int Wo, int R, int S, int stride, int pad,
                     int transposed) {
    if (N <= 0 || Hg <= 0 || Wg <= 0 || Cg <= 0 || Ho <= 0 || Wo <= 0 || R <= 0 || S <= 0 || stride <= 0 || pad < 0)
        return DSR_ERR_ARG;
    g->N = N; g->Hg = Hg; g->Wg = Wg; g->Cg = Cg; g->Ho = Ho; g->Wo = Wo;
    g->R = R; g->S = S; g->stride = stride; g->pad = pad; g->transposed = transposed;
    return DSR_OK;
}

extern "C" int dsr_conv_simt(const float* G, const float* Wk, const float* bias, float* out, int N, int Hg, int Wg,
                             int Cg, int Ho, int Wo, int Co, int R, int S, int stride, int pad, int transposed,
                             int act_out, void* stream) {
    DSR_REQUIRE(G && Wk && out && Co > 0, "bad arguments");
    Geom g;
    if (fill_geom(&g, N, Hg, Wg, Cg, Ho, Wo, R, S, stride, pad, transposed)) { dsr_set_error("conv_simt: bad geometry"); return DSR_ERR_ARG; }
    long M = (long)N * Ho * Wo;
    dim3 grid(dsr_cdiv(M, BM), dsr_cdiv(Co, BN));
    bool vec = (Cg % 16 == 0) && !((uintptr_t)G & 15);
    if (vec) conv_simt_kernel<true><<<grid, 256, 0, ST(stream)>>>(G, Wk, bias, out, g, Co, act_out);
    else conv_simt_kernel<false><<<grid, 256, 0, ST(stream)>>>(G, Wk, bias, out, g, Co, act_out);
    return dsr_check_launch("conv_simt");
}

extern "C" int dsr_wgrad_simt(const float* G, const float* D, float* dWk, int N, int Hg, int Wg, int Cg, int Ho,
                              int Wo, int Cd, int R, int S, int stride, int pad, void* stream) {
    DSR_REQUIRE(G && D && dWk && Cd > 0, "bad arguments");
    Geom g;
    if (fill_geom(&g, N, Hg, Wg, Cg, Ho, Wo, R, S, stride, pad, 0)) { dsr_set_error("wgrad_simt: bad geometry"); return DSR_ERR_ARG; }
    long M = (long)N * Ho * Wo;
    int K = R * S * Cg;
    int tk = dsr_cdiv(K, BM), tn = dsr_cdiv(Cd, BN);
    long splits = ((long)dsr_num_sms() * 4 + (long)tk * tn - 1) / ((long)tk * tn);
    long max_splits = (M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long mps = ((M + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (M + mps - 1) / mps;
    if (cudaMemsetAsync(dWk, 0, (size_t)K * Cd * sizeof(float), ST(stream)) != cudaSuccess) {
        dsr_set_error("wgrad_simt: memset failed"); return DSR_ERR_CUDA;
    }
    dim3 grid(tk, tn, (unsigned)splits);
    wgrad_simt_kernel<<<grid, 256, 0, ST(stream)>>>(G, D, dWk, g, Cd, mps);
    return dsr_check_launch("wgrad_simt");
}
