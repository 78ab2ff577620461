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

// 4x4 float64 linear solve helpers (device): solve X * M = R for X (rows of R independent), i.e.
// X = R * inv(M), by Gauss-Jordan with partial pivoting on M^T.  Returns false if singular.
__device__ inline bool pm_solve_right_4x4(const double M[16], const double *R, int nrows, double *X,
                                          double rel_tol) {
    // Solve M^T x_r^T = R_r^T for every row r.  Augmented matrix [M^T | R^T].
    double a[4][4 + 4];
    double scale = 0.0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            a[i][j] = M[j * 4 + i];
            scale = fmax(scale, fabs(a[i][j]));
        }
    for (int i = 0; i < 4; ++i)
        for (int r = 0; r < nrows; ++r) a[i][4 + r] = R[r * 4 + i];
    if (!(scale > 0.0) || !isfinite(scale)) return false;
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        double best = fabs(a[c][c]);
        for (int i = c + 1; i < 4; ++i)
            if (fabs(a[i][c]) > best) { best = fabs(a[i][c]); piv = i; }
        if (!(best > rel_tol * scale)) return false;
        if (piv != c)
            for (int j = 0; j < 4 + nrows; ++j) { double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
        const double inv = 1.0 / a[c][c];
        for (int j = c; j < 4 + nrows; ++j) a[c][j] *= inv;
        for (int i = 0; i < 4; ++i) {
            if (i == c) continue;
            const double f = a[i][c];
            if (f != 0.0)
                for (int j = c; j < 4 + nrows; ++j) a[i][j] -= f * a[c][j];
        }
    }
    for (int r = 0; r < nrows; ++r)
        for (int i = 0; i < 4; ++i) X[r * 4 + i] = a[i][4 + r];
    return true;
}
