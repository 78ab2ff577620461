// pm_chi2.cu — K3: N1 x N2 chi^2 histogram-distance cost matrix, register-tiled packed FP32.
//
// Reference: platymatch/estimate_transform/shape_context.py:88-99 (get_unary_distance) evaluated for
// every (moving, fixed) pair by the double loops at platymatch/_dock_widget.py:547-602.
//   cost[i][j] = 0.5 * sum_k (a_ik - b_jk)^2 / (a_ik + b_jk),   bins with a == b skipped.
// The metric is not bilinear (one divide per bin pair), so it does not reduce to a dot product and
// there is no tensor-core formulation; it runs on the FP32 pipes + MUFU.RCP.
//
// Three things make it fast:
//  (1) structural zeros.  Shape-context histograms of shell-like clouds populate ~190 of the 360
//      bins at all (r and theta are tied on a shell).  pm_chi2_operand records, per 128-nucleus block
//      and bin, whether any histogram of the block is non-empty (12 x u32 per block).  A CTA
//      (128 x 128 outputs) ORs / ANDs the masks of its row and column block: bins empty on both
//      sides are skipped (they contribute exactly 0), bins empty on ONE side contribute the other
//      side's value (a^2/a = a) and are folded into a per-row / per-column sum, and only bins
//      populated on both sides go through the full evaluation.
//  (2) two bins per reciprocal: t1/s1 + t2/s2 = (t1 s2 + t2 s1) / (s1 s2) halves the MUFU.RCP count
//      that bound the first version of this kernel (one MUFU per bin pair = 16/clk/SM).
//  (3) packed FP32 (FADD2/FMUL2/FFMA2, sm_100): two adjacent columns per instruction with the row
//      value as a scalar-broadcast operand -> 5 packed + 1 MUFU issue slots per 2 columns x 2 bins.
// Per (row, column, bin pair): 10 FP32 lane-operations + 1 MUFU, against 2 x 5 = 10 algorithmic FLOPs.
//
// Empty bins: both operands carry PM_CHI2_EPS (2^-60) instead of 0.  Then a == b == eps gives
// d = 0 exactly (contribution exactly 0, like the reference's "skip equal bins"), s = 2 eps and
// s1 * s2 >= 2^-118 stays a normal float, so the reciprocal is finite and no branch is needed;
// against a populated bin eps vanishes in the rounding (eps << ulp of any count / total).
// Identical histograms therefore give exactly 0.
#include <type_traits>
#include "pm_common.cuh"

#define PM_X2_TILE 128
#define PM_X2_KC 8           // bins per pipeline stage (4 pairs)
#define PM_X2_THREADS 256
#define PM_X2_NULL_BIN PM_NBINS   // operand row 360: all eps, pads the bin list to whole stages
#define PM_X2_MASK_WORDS 12

typedef unsigned long long pm_f32x2;   // two floats in one 64-bit register pair: {lo, hi}

__device__ __forceinline__ pm_f32x2 pm_pack(float lo, float hi) {
    pm_f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void pm_unpack(pm_f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ pm_f32x2 pm_add2(pm_f32x2 a, pm_f32x2 b) {
    pm_f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ pm_f32x2 pm_sub2(pm_f32x2 a, pm_f32x2 b) {
    pm_f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ pm_f32x2 pm_mul2(pm_f32x2 a, pm_f32x2 b) {
    pm_f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ pm_f32x2 pm_fma2(pm_f32x2 a, pm_f32x2 b, pm_f32x2 c) {
    pm_f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float pm_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ pm_f32x2 pm_rcp2(pm_f32x2 p) {
    float lo, hi;
    pm_unpack(p, lo, hi);
    return pm_pack(pm_rcp(lo), pm_rcp(hi));
}

__device__ __forceinline__ void pm_cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void pm_cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void pm_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// ------------------------------------------------------------------------------------ operand
// counts [n][360] u32 -> out [361][ld] float32 bin-major (count / row total, empty -> eps, pad columns
// and the null row 360 -> eps) and mask [ld/128][12]: bit k of a block = some histogram of the
// block has a non-empty bin k.  grid (ld/128), block 256; coalesced through a shared-memory transpose.
// T = uint32_t: integer counts, normalised here by the row total; T = float: values used as given.
template <typename T>
__global__ void __launch_bounds__(256) pm_chi2_operand_kernel(const T *__restrict__ counts, int n,
                                                              float *__restrict__ out, int ld,
                                                              uint32_t *__restrict__ mask) {
    constexpr bool kCounts = sizeof(T) == sizeof(uint32_t) && !std::is_floating_point<T>::value;
    __shared__ float tile[32][129];
    __shared__ float total[128];
    __shared__ uint32_t s_mask[PM_X2_MASK_WORDS];
    const int i0 = blockIdx.x * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;   // 8 warps
    if (threadIdx.x < PM_X2_MASK_WORDS) s_mask[threadIdx.x] = 0u;
    for (int r = warp; r < 128 && kCounts; r += 8) {        // row totals: one warp per histogram
        const int i = i0 + r;
        uint32_t s = 0;
        if (i < n)
            for (int k = lane; k < PM_NBINS; k += 32) s += (uint32_t)counts[(size_t)i * PM_NBINS + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) total[r] = (float)s;
    }
    __syncthreads();
    for (int k0 = 0; k0 < PM_NBINS; k0 += 32) {
        // load 128 histograms x 32 bins (lane = bin: 128-byte rows), mark non-empty bins
        uint32_t any = 0u;
        for (int r = warp; r < 128; r += 8) {
            const int i = i0 + r, k = k0 + lane;
            float v = PM_CHI2_EPS;
            if (i < n && k < PM_NBINS) {
                const T c = counts[(size_t)i * PM_NBINS + k];
                if (c > (T)0) { v = kCounts ? (float)c / total[r] : (float)c; any = 1u; }   // count / total, one rounding
            }
            tile[lane][r] = v;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, any != 0u);   // bit = bin within this group of 32
        if (lane == 0 && word) atomicOr(&s_mask[k0 >> 5], word);
        __syncthreads();
        for (int kk = warp; kk < 32; kk += 8) {
            const int k = k0 + kk;
            if (k < PM_NBINS)
                for (int c = lane; c < 128; c += 32) out[(size_t)k * ld + i0 + c] = tile[kk][c];
        }
        __syncthreads();
    }
    for (int c = threadIdx.x; c < 128; c += 256) out[(size_t)PM_X2_NULL_BIN * ld + i0 + c] = PM_CHI2_EPS;
    if (threadIdx.x < PM_X2_MASK_WORDS) mask[(size_t)blockIdx.x * PM_X2_MASK_WORDS + threadIdx.x] = s_mask[threadIdx.x];
}

extern "C" int pm_chi2_operand(const uint32_t *counts, int n, float *out, int ld, uint32_t *mask, void *stream) {
    PM_REQUIRE(counts && out && mask, "null pointer");
    PM_REQUIRE(n >= 1 && ld >= n && ld % PM_X2_TILE == 0, "need n >= 1 and ld = n rounded up to 128");
    pm_chi2_operand_kernel<uint32_t><<<ld / PM_X2_TILE, 256, 0, pm_stream(stream)>>>(counts, n, out, ld, mask);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

extern "C" int pm_chi2_operand_f32(const float *hist, int n, float *out, int ld, uint32_t *mask, void *stream) {
    PM_REQUIRE(hist && out && mask, "null pointer");
    PM_REQUIRE(n >= 1 && ld >= n && ld % PM_X2_TILE == 0, "need n >= 1 and ld = n rounded up to 128");
    pm_chi2_operand_kernel<float><<<ld / PM_X2_TILE, 256, 0, pm_stream(stream)>>>(hist, n, out, ld, mask);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

// ------------------------------------------------------------------------------------ cost
// Tiling: 128 x 128 outputs per CTA, 256 threads as 16 x 16, each thread an 8 x 8 register tile
// (rows ty*4..+3 and 64+ty*4..+3, columns tx*4..+3 and 64+tx*4..+3: LDS.128 operand reads without
// bank conflicts), accumulators as 8 x 4 packed column pairs.  Operand rows of the selected bins are
// staged with 16-byte cp.async into a double-buffered ring of 8-bin stages.
__global__ void __launch_bounds__(PM_X2_THREADS, 2)
pm_chi2_kernel(const float *__restrict__ a_t, int lda, const uint32_t *__restrict__ a_mask,
               const float *__restrict__ b_t, int ldb, const uint32_t *__restrict__ b_mask, int n2,
               int row_begin, int row_end, float *__restrict__ cost, int ldc) {
    __shared__ __align__(16) float As[2][PM_X2_KC][PM_X2_TILE];
    __shared__ __align__(16) float Bs[2][PM_X2_KC][PM_X2_TILE];
    __shared__ float s_rowsum[PM_X2_TILE], s_colsum[PM_X2_TILE];
    __shared__ unsigned short s_both[PM_NBINS + PM_X2_KC], s_single[PM_NBINS];
    __shared__ uint32_t s_wa[PM_X2_MASK_WORDS], s_wb[PM_X2_MASK_WORDS];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = row_begin + blockIdx.y * PM_X2_TILE, j0 = blockIdx.x * PM_X2_TILE;

    if (tid < PM_X2_MASK_WORDS) {
        s_wa[tid] = a_mask[(size_t)(i0 / PM_X2_TILE) * PM_X2_MASK_WORDS + tid];
        s_wb[tid] = b_mask[(size_t)blockIdx.x * PM_X2_MASK_WORDS + tid];
    }
    __syncthreads();
    // ordered bin lists: both sides populated / only the row side / only the column side
    int n_both = 0, n_aonly = 0, n_bonly = 0;
#pragma unroll
    for (int w = 0; w < PM_X2_MASK_WORDS; ++w) {
        n_both += __popc(s_wa[w] & s_wb[w]);
        n_aonly += __popc(s_wa[w] & ~s_wb[w]);
        n_bonly += __popc(~s_wa[w] & s_wb[w]);
    }
    for (int k = tid; k < PM_NBINS; k += PM_X2_THREADS) {
        const int w = k >> 5;
        const uint32_t bit = 1u << (k & 31), below = bit - 1u;
        const uint32_t wa = s_wa[w], wb = s_wb[w];
        int p_both = 0, p_a = 0, p_b = 0;
        for (int q = 0; q < w; ++q) {
            p_both += __popc(s_wa[q] & s_wb[q]);
            p_a += __popc(s_wa[q] & ~s_wb[q]);
            p_b += __popc(~s_wa[q] & s_wb[q]);
        }
        if (wa & wb & bit) s_both[p_both + __popc(wa & wb & below)] = (unsigned short)k;
        else if (wa & ~wb & bit) s_single[p_a + __popc(wa & ~wb & below)] = (unsigned short)k;
        else if (~wa & wb & bit) s_single[n_aonly + p_b + __popc(~wa & wb & below)] = (unsigned short)k;
    }
    const int n_stage = (n_both + PM_X2_KC - 1) / PM_X2_KC;
    if (tid < PM_X2_KC && n_both + tid < n_stage * PM_X2_KC) s_both[n_both + tid] = PM_X2_NULL_BIN;
    __syncthreads();

    // staging role: thread copies one float4 of A and one of B per stage
    const int lk = tid >> 5, lc = (tid & 31) * 4;
    const float *ag = a_t + i0 + lc;
    const float *bg = b_t + j0 + lc;
    if (n_stage > 0) {
        const int k = s_both[lk];
        pm_cp_async16(&As[0][lk][lc], ag + (size_t)k * lda);
        pm_cp_async16(&Bs[0][lk][lc], bg + (size_t)k * ldb);
    }
    pm_cp_async_commit();

    // one-sided bins: (a - b)^2 / (a + b) with b empty is a; eps is taken off so that a == eps adds 0
    if (tid < PM_X2_TILE) {
        float s = 0.0f;
        for (int q = 0; q < n_aonly; ++q) s += a_t[(size_t)s_single[q] * lda + i0 + tid] - PM_CHI2_EPS;
        s_rowsum[tid] = s;
    } else {
        float s = 0.0f;
        const int c = tid - PM_X2_TILE;
        for (int q = 0; q < n_bonly; ++q) s += b_t[(size_t)s_single[n_aonly + q] * ldb + j0 + c] - PM_CHI2_EPS;
        s_colsum[c] = s;
    }

    pm_f32x2 acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0ull;

    for (int st = 0; st < n_stage; ++st) {
        const int buf = st & 1;
        if (st + 1 < n_stage) {
            const int k = s_both[(st + 1) * PM_X2_KC + lk];
            pm_cp_async16(&As[buf ^ 1][lk][lc], ag + (size_t)k * lda);
            pm_cp_async16(&Bs[buf ^ 1][lk][lc], bg + (size_t)k * ldb);
            pm_cp_async_commit();
            pm_cp_async_wait<1>();
        } else {
            pm_cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll
        for (int kp = 0; kp < PM_X2_KC; kp += 2) {
            const float4 a1_lo = *reinterpret_cast<const float4 *>(&As[buf][kp][ty * 4]);
            const float4 a1_hi = *reinterpret_cast<const float4 *>(&As[buf][kp][64 + ty * 4]);
            const float4 a2_lo = *reinterpret_cast<const float4 *>(&As[buf][kp + 1][ty * 4]);
            const float4 a2_hi = *reinterpret_cast<const float4 *>(&As[buf][kp + 1][64 + ty * 4]);
            const ulonglong2 b1_lo = *reinterpret_cast<const ulonglong2 *>(&Bs[buf][kp][tx * 4]);
            const ulonglong2 b1_hi = *reinterpret_cast<const ulonglong2 *>(&Bs[buf][kp][64 + tx * 4]);
            const ulonglong2 b2_lo = *reinterpret_cast<const ulonglong2 *>(&Bs[buf][kp + 1][tx * 4]);
            const ulonglong2 b2_hi = *reinterpret_cast<const ulonglong2 *>(&Bs[buf][kp + 1][64 + tx * 4]);
            const float a1[8] = {a1_lo.x, a1_lo.y, a1_lo.z, a1_lo.w, a1_hi.x, a1_hi.y, a1_hi.z, a1_hi.w};
            const float a2[8] = {a2_lo.x, a2_lo.y, a2_lo.z, a2_lo.w, a2_hi.x, a2_hi.y, a2_hi.z, a2_hi.w};
            const pm_f32x2 b1[4] = {b1_lo.x, b1_lo.y, b1_hi.x, b1_hi.y};
            const pm_f32x2 b2[4] = {b2_lo.x, b2_lo.y, b2_hi.x, b2_hi.y};
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const pm_f32x2 A1 = pm_pack(a1[r], a1[r]), A2 = pm_pack(a2[r], a2[r]);   // scalar-broadcast operand
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const pm_f32x2 d1 = pm_sub2(A1, b1[c]), s1 = pm_add2(A1, b1[c]);
                    const pm_f32x2 d2 = pm_sub2(A2, b2[c]), s2 = pm_add2(A2, b2[c]);
                    const pm_f32x2 t1 = pm_mul2(d1, d1), t2 = pm_mul2(d2, d2);
                    const pm_f32x2 den = pm_mul2(s1, s2);
                    const pm_f32x2 num = pm_fma2(t2, s1, pm_mul2(t1, s2));
                    acc[r][c] = pm_fma2(num, pm_rcp2(den), acc[r][c]);
                }
            }
        }
        __syncthreads();
    }
    if (n_stage == 0) __syncthreads();   // s_rowsum / s_colsum visible

    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<size_t>(cost) & 15) == 0);
    float cs[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) cs[c] = s_colsum[(c < 4 ? 0 : 64) + tx * 4 + (c & 3)];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int lr = (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
        const int i = i0 + lr;
        if (i >= row_end) continue;
        const float rs = s_rowsum[lr];
        float *crow = cost + (size_t)(i - row_begin) * ldc;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = j0 + h * 64 + tx * 4;
            float v[4];
            pm_unpack(acc[r][h * 2 + 0], v[0], v[1]);
            pm_unpack(acc[r][h * 2 + 1], v[2], v[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = 0.5f * (v[e] + (rs + cs[h * 4 + e]));
            if (vec_ok && j + 3 < n2) {
                *reinterpret_cast<float4 *>(crow + j) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (j + e < n2) crow[j + e] = v[e];
            }
        }
    }
}

extern "C" int pm_chi2_cost(const float *a_t, int lda, const uint32_t *a_mask, int n1, const float *b_t, int ldb,
                            const uint32_t *b_mask, int n2, int row_begin, int row_end, float *cost, int ldc,
                            void *stream) {
    PM_REQUIRE(a_t && b_t && a_mask && b_mask && cost, "null pointer");
    PM_REQUIRE(n1 >= 1 && n2 >= 1, "empty matrix");
    PM_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= n1, "bad row range");
    PM_REQUIRE(ldc >= n2, "ldc < n2");
    // operand tiles are read without bounds checks: leading dimensions cover whole 128-blocks
    PM_REQUIRE(lda % PM_X2_TILE == 0 && ldb % PM_X2_TILE == 0 && lda >= n1 && ldb >= n2,
               "lda/ldb must be n1/n2 rounded up to 128 (pm_chi2_operand layout)");
    PM_REQUIRE((reinterpret_cast<size_t>(a_t) & 15) == 0 && (reinterpret_cast<size_t>(b_t) & 15) == 0,
               "operands must be 16-byte aligned");
    PM_REQUIRE(row_begin % PM_X2_TILE == 0, "row_begin must be a multiple of 128 (mask blocks)");
    const int rows = row_end - row_begin;
    if (rows == 0) return PM_OK;
    const int tiles_y = (rows + PM_X2_TILE - 1) / PM_X2_TILE;
    dim3 grid((n2 + PM_X2_TILE - 1) / PM_X2_TILE, tiles_y);
    pm_chi2_kernel<<<grid, PM_X2_THREADS, 0, pm_stream(stream)>>>(a_t, lda, a_mask, b_t, ldb, b_mask, n2, row_begin,
                                                                  row_end, cost, ldc);
    PM_LAUNCH_CHECK();
    return PM_OK;
}
