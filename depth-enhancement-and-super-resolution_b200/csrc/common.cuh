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
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
#endif
