// Common helpers for the dsr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define DSR_OK 0
#define DSR_ERR_ARG (-1)
#define DSR_ERR_CUDA (-2)
#define DSR_ERR_UNSUPPORTED (-3)

#ifndef DSR_HD
#ifdef __CUDACC__
#define DSR_HD __host__ __device__ __forceinline__
#else
#define DSR_HD inline
#endif
#endif

void dsr_set_error(const char* fmt, ...);
int dsr_check_launch(const char* what);
int dsr_num_sms();

#define DSR_REQUIRE(cond, msg)                                  \
    do {                                                        \
        if (!(cond)) {                                          \
            dsr_set_error("%s: %s", __func__, msg);             \
            return DSR_ERR_ARG;                                 \
        }                                                       \
    } while (0)

static inline int dsr_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// grid size for a grid-stride kernel: a multiple of the SM count, capped by the work available
static inline int dsr_grid(long long work_items, int threads, int ctas_per_sm = 8) {
    long long need = (work_items + threads - 1) / threads;
    long long cap = (long long)dsr_num_sms() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum, result valid in thread 0.  blockDim.x must be a multiple of 32 (<= 1024)
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem32) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? smem32[threadIdx.x] : T(0);
    if (w == 0) v = warp_sum(v);
    return v;
}
// Normalisation statistics -> (mean, scale, shift) of one (n, c): the arithmetic of dsr_norm_finalize, shared with the
// consumers that fold the finalisation into their own prologue (dsr_tc_prep_fin, dsr_norm_apply_fwd_fin) so that both routes
// are bit-identical.  groups 0: InstanceNorm2d(affine=False) models/networks.py:30; groups > 0: GroupNorm(groups, C, affine)
// models/translation_network.py:46.
struct NormFin {
    const double* sums;      // [N][C][2] (sum, sum of squares) over P pixels; NULL = no folded finalisation
    const float* gamma;      // [C] or NULL
    const float* beta;       // [C] or NULL
    float* prm_out;          // [3][N][C] written for later readers (backward pass) by one block per sample
    long P;
    int groups;
    float eps;
};
__device__ __forceinline__ void norm_fin_one(const double* __restrict__ sums, long P, int groups, const float* __restrict__ gamma,
                                             const float* __restrict__ beta, float eps, int n, int c, int C,
                                             float& mean_f, float& scale, float& shift) {
    double s = 0, q = 0, cnt;
    const long idx = (long)n * C + c;
    if (groups == 0) { s = sums[idx * 2]; q = sums[idx * 2 + 1]; cnt = (double)P; }
    else {
        int cg = C / groups, g0 = (c / cg) * cg;
        for (int k = 0; k < cg; ++k) { s += sums[((long)n * C + g0 + k) * 2]; q += sums[((long)n * C + g0 + k) * 2 + 1]; }
        cnt = (double)P * cg;
    }
    double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0) var = 0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    mean_f = (float)mean;
    scale = groups == 0 ? rstd : rstd * (gamma ? gamma[c] : 1.f);
    shift = groups == 0 ? 0.f : (beta ? beta[c] : 0.f);
}
// One element of the InstanceNorm2d(affine=False) [+ReLU] backward: dx = rstd * (g' - mean(g') - xhat * mean(g' * xhat)), g' = g
// gated by xhat > 0 under a fused ReLU.  Shared by dsr_in_bwd_apply (nn.cu) and the fused apply + operand preparation
// (dsr_tc_prep_in_bwd, conv_tc.cu) so that both routes produce the same bits.
__device__ __forceinline__ float in_bwd_one(float x, float g, float mean, float rstd, float m1, float m2, bool relu) {
    const float xh = (x - mean) * rstd;
    if (relu && !(xh > 0.f)) g = 0.f;
    return rstd * (g - m1 - xh * m2);
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
#endif
