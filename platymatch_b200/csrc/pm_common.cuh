// pm_common.cuh — shared helpers for libplatymatch_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/platymatch_b200.h"

void pm_set_error(const char *fmt, ...);
void pm_count_launches(int n);

#define PM_CUDA_TRY(expr)                                                                     \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            pm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return PM_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define PM_LAUNCH_CHECK_N(n)                                                                   \
    do {                                                                                      \
        pm_count_launches(n);                                                                 \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            pm_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return PM_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define PM_LAUNCH_CHECK() PM_LAUNCH_CHECK_N(1)

#define PM_REQUIRE(cond, msg)                                                  \
    do {                                                                       \
        if (!(cond)) {                                                         \
            pm_set_error("%s: invalid argument: %s", __func__, msg);          \
            return PM_ERR_INVALID_ARGUMENT;                                    \
        }                                                                      \
    } while (0)

static inline cudaStream_t pm_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ double pm_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide float64 sum (all threads get the result).  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double pm_block_sum(double v, double *smem32) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = pm_warp_sum(v);
    __syncthreads();
    if (lane == 0) smem32[warp] = v;
    __syncthreads();
    double r = (lane < nw) ? smem32[lane] : 0.0;
    r = pm_warp_sum(r);
    return r;
}

#include "pm_linalg.cuh"   // pm_solve_right_4x4, pm_jacobi_sym4, pm_affine_from_normal_eq, pm_similar_from_moments
