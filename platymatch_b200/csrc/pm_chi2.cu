// pm_chi2.cu — K3: N1 x N2 chi^2 histogram-distance cost matrix, register-tiled FP32.
//
// Reference: platymatch/estimate_transform/shape_context.py:88-99 (get_unary_distance) evaluated for
// every (moving, fixed) pair by the double loops at platymatch/_dock_widget.py:547-602.
//   cost[i][j] = 0.5 * sum_k (a_ik - b_jk)^2 / (a_ik + b_jk),   bins with a == b skipped.
// The metric is not bilinear (one divide per bin pair), so it does not reduce to a dot product and
// there is no tensor-core formulation; it runs on the FP32 pipes + MUFU.RCP.
//
// Tiling: 128 x 128 outputs per CTA, 256 threads as 16 x 16, each thread an 8 x 8 register tile
// (split 4+4 along both axes so the LDS.128 operand reads are bank-conflict free).  Operands are
// bin-major ([360][ld], written by pm_normalise_hist) so a K-chunk of a tile is KC rows of 128
// contiguous floats: staged with 16-byte cp.async into a double-buffered shared-memory ring.
// Per bin pair: FADD d=a-b, FADD s=a+b, MUFU.RCP r=1/s, FMUL t=d*d, FFMA acc+=t*r.
// Empty bins: A carries exact zeros, B carries PM_CHI2_ZERO_SENTINEL (1e-30) for zeros, so
// 0/0 bins give t = 1e-60 -> flushed to 0 and r = 1e30 finite: contribution exactly 0, no branch.
#include "pm_common.cuh"

#define PM_X2_TILE 128
#define PM_X2_KC 8
#define PM_X2_THREADS 256

__device__ __forceinline__ float pm_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void pm_cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void pm_cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void pm_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__global__ void __launch_bounds__(PM_X2_THREADS, 2)
pm_chi2_kernel(const float *__restrict__ a_t, int lda, const float *__restrict__ b_t, int ldb, int n2,
               int row_begin, int row_end, float *__restrict__ cost, int ldc) {
    __shared__ __align__(16) float As[2][PM_X2_KC][PM_X2_TILE];
    __shared__ __align__(16) float Bs[2][PM_X2_KC][PM_X2_TILE];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = row_begin + blockIdx.y * PM_X2_TILE, j0 = blockIdx.x * PM_X2_TILE;
    // staging role: thread copies one float4 of A and one of B per K-chunk
    const int lk = tid >> 5, lc = (tid & 31) * 4;
    const float *ag = a_t + (size_t)lk * lda + i0 + lc;
    const float *bg = b_t + (size_t)lk * ldb + j0 + lc;

    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;

    pm_cp_async16(&As[0][lk][lc], ag);
    pm_cp_async16(&Bs[0][lk][lc], bg);
    pm_cp_async_commit();
    constexpr int NCHUNK = PM_NBINS / PM_X2_KC;
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < NCHUNK) {
            pm_cp_async16(&As[buf ^ 1][lk][lc], ag + (size_t)(ch + 1) * PM_X2_KC * lda);
            pm_cp_async16(&Bs[buf ^ 1][lk][lc], bg + (size_t)(ch + 1) * PM_X2_KC * ldb);
            pm_cp_async_commit();
            pm_cp_async_wait<1>();
        } else {
            pm_cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PM_X2_KC; ++k) {
            const float4 a_lo = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
            const float4 a_hi = *reinterpret_cast<const float4 *>(&As[buf][k][64 + ty * 4]);
            const float4 b_lo = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
            const float4 b_hi = *reinterpret_cast<const float4 *>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
            const float b[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float d = a[r] - b[c];
                    const float s = a[r] + b[c];
                    const float t = d * d;
                    acc[r][c] = fmaf(t, pm_rcp(s), acc[r][c]);
                }
        }
        __syncthreads();
    }

    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<size_t>(cost) & 15) == 0);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
        if (i >= row_end) continue;
        float *crow = cost + (size_t)(i - row_begin) * ldc;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = j0 + h * 64 + tx * 4;
            const float v0 = 0.5f * acc[r][h * 4 + 0], v1 = 0.5f * acc[r][h * 4 + 1],
                        v2 = 0.5f * acc[r][h * 4 + 2], v3 = 0.5f * acc[r][h * 4 + 3];
            if (vec_ok && j + 3 < n2) {
                *reinterpret_cast<float4 *>(crow + j) = make_float4(v0, v1, v2, v3);
            } else {
                if (j < n2) crow[j] = v0;
                if (j + 1 < n2) crow[j + 1] = v1;
                if (j + 2 < n2) crow[j + 2] = v2;
                if (j + 3 < n2) crow[j + 3] = v3;
            }
        }
    }
}

extern "C" int pm_chi2_cost(const float *a_t, int lda, int n1, const float *b_t, int ldb, int n2, int row_begin,
                            int row_end, float *cost, int ldc, void *stream) {
    PM_REQUIRE(a_t && b_t && cost, "null pointer");
    PM_REQUIRE(n1 >= 1 && n2 >= 1, "empty matrix");
    PM_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= n1, "bad row range");
    PM_REQUIRE(ldc >= n2, "ldc < n2");
    // operand tiles are read without bounds checks: leading dimensions must cover whole tiles
    const int need_a = ((n1 + PM_X2_TILE - 1) / PM_X2_TILE) * PM_X2_TILE;
    const int need_b = ((n2 + PM_X2_TILE - 1) / PM_X2_TILE) * PM_X2_TILE;
    PM_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, "lda/ldb must be multiples of 4");
    PM_REQUIRE((reinterpret_cast<size_t>(a_t) & 15) == 0 && (reinterpret_cast<size_t>(b_t) & 15) == 0,
               "operands must be 16-byte aligned");
    PM_REQUIRE(row_begin % 4 == 0, "row_begin must be a multiple of 4");
    // a tile starting at row_begin may run up to row_begin + k*128 <= lda
    const int rows = row_end - row_begin;
    if (rows == 0) return PM_OK;
    const int tiles_y = (rows + PM_X2_TILE - 1) / PM_X2_TILE;
    PM_REQUIRE(row_begin + tiles_y * PM_X2_TILE <= lda || need_a <= lda, "lda must cover whole 128-row tiles");
    PM_REQUIRE(row_begin + tiles_y * PM_X2_TILE <= lda, "lda must cover the last 128-row tile of the range");
    PM_REQUIRE(need_b <= ldb, "ldb must be n2 rounded up to 128");
    dim3 grid((n2 + PM_X2_TILE - 1) / PM_X2_TILE, tiles_y);
    pm_chi2_kernel<<<grid, PM_X2_THREADS, 0, pm_stream(stream)>>>(a_t, lda, b_t, ldb, n2, row_begin, row_end, cost,
                                                                  ldc);
    PM_LAUNCH_CHECK();
    return PM_OK;
}
