// pm_cloud.cu — K0 cloud statistics (centroid + PCA first axis), K1 mean pairwise distance, and the
// small 4x4 / per-point operations of the path.  float64 throughout (the reference is float64 and the
// final transform must hold 1e-4 on translations of ~1e2..1e3 px).
//
// Reference: platymatch/utils/utils.py:48-56 (get_centroid), :58-75 (get_mean_distance);
// platymatch/estimate_transform/shape_context.py:162-165 (PCA axis); find_transform.py:4-17
// (get_affine_transform); apply_transform.py:3-17 (apply_affine_transform).
#include "pm_common.cuh"

// ------------------------------------------------------------------------------------------- K0
// One CTA: two-pass mean / centred covariance, then a cyclic-Jacobi 3x3 eigen-decomposition in
// thread 0.  N <= ~1e5 points of 24 B: a single SM streams that in a few microseconds.
__global__ void __launch_bounds__(1024) pm_cloud_stats_kernel(const double *__restrict__ pts, int n,
                                                              double *__restrict__ stats) {
    __shared__ double red[32];
    __shared__ double mean_s[3];
    double s0 = 0, s1 = 0, s2 = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s0 += pts[3 * i];
        s1 += pts[3 * i + 1];
        s2 += pts[3 * i + 2];
    }
    s0 = pm_block_sum(s0, red);
    s1 = pm_block_sum(s1, red);
    s2 = pm_block_sum(s2, red);
    if (threadIdx.x == 0) {
        mean_s[0] = s0 / n;
        mean_s[1] = s1 / n;
        mean_s[2] = s2 / n;
    }
    __syncthreads();
    const double m0 = mean_s[0], m1 = mean_s[1], m2 = mean_s[2];
    double c[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double a = pts[3 * i] - m0, b = pts[3 * i + 1] - m1, d = pts[3 * i + 2] - m2;
        c[0] += a * a; c[1] += a * b; c[2] += a * d;
        c[3] += b * b; c[4] += b * d; c[5] += d * d;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] = pm_block_sum(c[k], red);
    if (threadIdx.x != 0) return;
    const double inv = 1.0 / (double)(n - 1);
    double A[3][3] = {{c[0] * inv, c[1] * inv, c[2] * inv},
                      {c[1] * inv, c[3] * inv, c[4] * inv},
                      {c[2] * inv, c[4] * inv, c[5] * inv}};
    stats[0] = m0; stats[1] = m1; stats[2] = m2;
    stats[6] = A[0][0]; stats[7] = A[0][1]; stats[8] = A[0][2];
    stats[9] = A[1][1]; stats[10] = A[1][2]; stats[11] = A[2][2];
    stats[12] = (double)n;
    // cyclic Jacobi: A <- J^T A J, V <- V J
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-18 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; ++k) {  // columns p,q
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = cs * akp - sn * akq;
                    A[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; ++k) {  // rows p,q
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = cs * apk - sn * aqk;
                    A[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = cs * vkp - sn * vkq;
                    V[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    int order[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (A[order[j]][order[j]] > A[order[i]][order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
    const int e = order[0];
    double x0[3] = {V[0][e], V[1][e], V[2][e]};
    const double nrm = sqrt(x0[0] * x0[0] + x0[1] * x0[1] + x0[2] * x0[2]);
    int am = 0;
    for (int k = 1; k < 3; ++k)
        if (fabs(x0[k]) > fabs(x0[am])) am = k;
    const double sgn = (x0[am] < 0 ? -1.0 : 1.0) / nrm;  // svd_flip: largest-|.| entry positive
    stats[3] = x0[0] * sgn; stats[4] = x0[1] * sgn; stats[5] = x0[2] * sgn;
    stats[13] = A[order[0]][order[0]]; stats[14] = A[order[1]][order[1]]; stats[15] = A[order[2]][order[2]];
}

extern "C" int pm_cloud_stats(const double *pts, int n, double *stats, void *stream) {
    PM_REQUIRE(pts && stats, "null pointer");
    PM_REQUIRE(n >= 2, "need at least 2 points");
    pm_cloud_stats_kernel<<<1, 1024, 0, pm_stream(stream)>>>(pts, n, stats);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

// ------------------------------------------------------------------------------------------- K1
// Upper-triangular tile pairs; each CTA stages the column tile in shared memory, every thread owns
// one row point.  Partials are written per tile pair and reduced in a fixed order (deterministic).
#define PM_MD_TILE 256

__global__ void __launch_bounds__(PM_MD_TILE) pm_mean_distance_tiles(const double *__restrict__ pts, int n,
                                                                     int tiles, double *__restrict__ partial) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;  // partial[] pre-zeroed
    __shared__ double sx[PM_MD_TILE], sy[PM_MD_TILE], sz[PM_MD_TILE];
    __shared__ double red[32];
    const int j0 = bj * PM_MD_TILE;
    const int jn = min(PM_MD_TILE, n - j0);
    if ((int)threadIdx.x < jn) {
        sx[threadIdx.x] = pts[3 * (j0 + threadIdx.x)];
        sy[threadIdx.x] = pts[3 * (j0 + threadIdx.x) + 1];
        sz[threadIdx.x] = pts[3 * (j0 + threadIdx.x) + 2];
    }
    __syncthreads();
    const int i = bi * PM_MD_TILE + threadIdx.x;
    double acc = 0.0;
    if (i < n) {
        const double px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
        const int jstart = (bi == bj) ? (int)threadIdx.x + 1 : 0;
        double acc2 = 0.0;
        int j = jstart;
        for (; j + 1 < jn; j += 2) {
            const double a0 = px - sx[j], a1 = py - sy[j], a2 = pz - sz[j];
            const double b0 = px - sx[j + 1], b1 = py - sy[j + 1], b2 = pz - sz[j + 1];
            acc += sqrt(a0 * a0 + a1 * a1 + a2 * a2);
            acc2 += sqrt(b0 * b0 + b1 * b1 + b2 * b2);
        }
        if (j < jn) {
            const double a0 = px - sx[j], a1 = py - sy[j], a2 = pz - sz[j];
            acc += sqrt(a0 * a0 + a1 * a1 + a2 * a2);
        }
        acc += acc2;
    }
    acc = pm_block_sum(acc, red);
    if (threadIdx.x == 0) partial[bi * tiles + bj] = acc;
}

__global__ void __launch_bounds__(1024) pm_mean_distance_final(const double *__restrict__ partial, int count,
                                                               int n, double *__restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += partial[i];
    s = pm_block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s / (0.5 * (double)n * (double)(n - 1));
}

extern "C" size_t pm_mean_distance_workspace_bytes(int n) {
    if (n < 1) return 0;
    const size_t t = (size_t)((n + PM_MD_TILE - 1) / PM_MD_TILE);
    return t * t * sizeof(double);
}

extern "C" int pm_mean_distance(const double *pts, int n, double *out_mean, void *workspace,
                                size_t workspace_bytes, void *stream) {
    PM_REQUIRE(pts && out_mean && workspace, "null pointer");
    PM_REQUIRE(n >= 2, "need at least 2 points");
    if (workspace_bytes < pm_mean_distance_workspace_bytes(n)) {
        pm_set_error("pm_mean_distance: workspace too small");
        return PM_ERR_WORKSPACE;
    }
    const int tiles = (n + PM_MD_TILE - 1) / PM_MD_TILE;
    cudaStream_t s = pm_stream(stream);
    PM_CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)tiles * tiles * sizeof(double), s));
    pm_mean_distance_tiles<<<dim3(tiles, tiles), PM_MD_TILE, 0, s>>>(pts, n, tiles, (double *)workspace);
    PM_LAUNCH_CHECK();
    pm_mean_distance_final<<<1, 1024, 0, s>>>((const double *)workspace, tiles * tiles, n, out_mean);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

// ------------------------------------------------------------------------------------ small ops
// Least-squares affine over K pairs: A = F_h M_h^T (M_h M_h^T)^+ == fixed_h @ pinv(moving_h), also for
// rank-deficient (coplanar, collinear, K < 4) point sets, where it is pinv's minimum-norm solution
// (pm_linalg.cuh).  Moving coordinates are shifted by moving[0] before forming the normal equations
// (block conditioning), and the shift is folded back into the translation column.
__global__ void __launch_bounds__(256) pm_fit_affine_kernel(const double *__restrict__ moving,
                                                            const double *__restrict__ fixed, int k,
                                                            double *__restrict__ A) {
    __shared__ double red[32];
    const double c0 = moving[0], c1 = moving[1], c2 = moving[2];
    double mm[10] = {0}, fm[12] = {0};
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const double m[4] = {moving[3 * i] - c0, moving[3 * i + 1] - c1, moving[3 * i + 2] - c2, 1.0};
        const double f[3] = {fixed[3 * i], fixed[3 * i + 1], fixed[3 * i + 2]};
        int q = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b) mm[q++] += m[a] * m[b];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) fm[a * 4 + b] += f[a] * m[b];
    }
#pragma unroll
    for (int q = 0; q < 10; ++q) mm[q] = pm_block_sum(mm[q], red);
#pragma unroll
    for (int q = 0; q < 12; ++q) fm[q] = pm_block_sum(fm[q], red);
    if (threadIdx.x != 0) return;
    double M[16];
    int q = 0;
    for (int a = 0; a < 4; ++a)
        for (int b = a; b < 4; ++b) { M[a * 4 + b] = mm[q]; M[b * 4 + a] = mm[q]; ++q; }
    const double shift[3] = {c0, c1, c2};
    pm_affine_from_normal_eq(M, fm, shift, 1e-14, A);     // rank-deficient point sets: pinv's minimum-norm answer
}

extern "C" int pm_fit_affine(const double *moving, const double *fixed, int k, double *A, void *stream) {
    PM_REQUIRE(moving && fixed && A, "null pointer");
    PM_REQUIRE(k >= 1, "need at least 1 pair");
    pm_fit_affine_kernel<<<1, 256, 0, pm_stream(stream)>>>(moving, fixed, k, A);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

// get_similar_transform (find_transform.py:21-99): Horn's closed form over K pairs, two passes like the reference
// (centroids :27-28, then centred moments :31-54,:88-94).  One CTA, fixed-order reductions.
__global__ void __launch_bounds__(256) pm_fit_similar_kernel(const double *__restrict__ moving,
                                                             const double *__restrict__ fixed, int k,
                                                             double *__restrict__ A) {
    __shared__ double red[32];
    double c[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < k; i += blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) { c[a] += moving[3 * i + a]; c[3 + a] += fixed[3 * i + a]; }
#pragma unroll
    for (int q = 0; q < 6; ++q) c[q] = pm_block_sum(c[q], red) / (double)k;
    double v[11] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const double p[3] = {moving[3 * i] - c[0], moving[3 * i + 1] - c[1], moving[3 * i + 2] - c[2]};
        const double y[3] = {fixed[3 * i] - c[3], fixed[3 * i + 1] - c[4], fixed[3 * i + 2] - c[5]};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int b = 0; b < 3; ++b) v[a * 3 + b] += p[a] * y[b];
            v[9] += p[a] * p[a];
            v[10] += y[a] * y[a];
        }
    }
#pragma unroll
    for (int q = 0; q < 11; ++q) v[q] = pm_block_sum(v[q], red);
    if (threadIdx.x != 0) return;
    double out[16];
    pm_similar_from_moments(c, c + 3, v, v[9], v[10], out);
    for (int e = 0; e < 16; ++e) A[e] = out[e];
}

extern "C" int pm_fit_similar(const double *moving, const double *fixed, int k, double *A, void *stream) {
    PM_REQUIRE(moving && fixed && A, "null pointer");
    PM_REQUIRE(k >= 1, "need at least 1 pair");
    pm_fit_similar_kernel<<<1, 256, 0, pm_stream(stream)>>>(moving, fixed, k, A);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

// np.argmax over the hypotheses' inlier counts (first maximum, _dock_widget.py:683-703) on the device, so that
// the registration never returns to the host between RANSAC and ICP: best index + that hypothesis' 4x4.
__global__ void pm_select_best_kernel(const int32_t *__restrict__ inliers, const double *__restrict__ A_all, int n_hyp,
                                      int32_t *__restrict__ best, double *__restrict__ A_best) {
    __shared__ int s_best;
    if (threadIdx.x == 0) {
        int b = 0;
        for (int q = 1; q < n_hyp; ++q)
            if (inliers[q] > inliers[b]) b = q;
        s_best = b;
        best[0] = b;
    }
    __syncthreads();
    if (threadIdx.x < 16) A_best[threadIdx.x] = A_all[(size_t)s_best * 16 + threadIdx.x];
}

extern "C" int pm_select_best(const int32_t *inliers, const double *A_all, int n_hyp, int32_t *best, double *A_best,
                              void *stream) {
    PM_REQUIRE(inliers && A_all && best && A_best, "null pointer");
    PM_REQUIRE(n_hyp >= 1, "need at least one hypothesis");
    pm_select_best_kernel<<<1, 32, 0, pm_stream(stream)>>>(inliers, A_all, n_hyp, best, A_best);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

__global__ void pm_apply_affine_kernel(const double *__restrict__ pts, int n, const double *__restrict__ A,
                                       double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    // same association as np.matmul's dot over [x, y, z, 1]
    out[3 * i] = ((A[0] * x + A[1] * y) + A[2] * z) + A[3];
    out[3 * i + 1] = ((A[4] * x + A[5] * y) + A[6] * z) + A[7];
    out[3 * i + 2] = ((A[8] * x + A[9] * y) + A[10] * z) + A[11];
}

extern "C" int pm_apply_affine(const double *pts, int n, const double *A, double *out, void *stream) {
    PM_REQUIRE(pts && A && out, "null pointer");
    PM_REQUIRE(n >= 0, "negative size");
    if (n == 0) return PM_OK;
    pm_apply_affine_kernel<<<(n + 255) / 256, 256, 0, pm_stream(stream)>>>(pts, n, A, out);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

__global__ void pm_gather_points_kernel(const double *__restrict__ pts, const int32_t *__restrict__ index, int k,
                                        double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const int s = index[i];
    out[3 * i] = pts[3 * s];
    out[3 * i + 1] = pts[3 * s + 1];
    out[3 * i + 2] = pts[3 * s + 2];
}

extern "C" int pm_gather_points(const double *pts, const int32_t *index, int k, double *out, void *stream) {
    PM_REQUIRE(pts && index && out, "null pointer");
    if (k <= 0) return PM_OK;
    pm_gather_points_kernel<<<(k + 255) / 256, 256, 0, pm_stream(stream)>>>(pts, index, k, out);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

__global__ void pm_compose_kernel(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C) {
    const int r = threadIdx.x >> 2, c = threadIdx.x & 3;
    double s = 0.0;
    for (int k = 0; k < 4; ++k) s += A[r * 4 + k] * B[k * 4 + c];
    C[r * 4 + c] = s;
}

extern "C" int pm_compose(const double *A, const double *B, double *C, void *stream) {
    PM_REQUIRE(A && B && C, "null pointer");
    PM_REQUIRE(C != A && C != B, "output must not alias inputs");
    pm_compose_kernel<<<1, 16, 0, pm_stream(stream)>>>(A, B, C);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

// ---- Euclidean distance matrix (scipy.spatial.distance.cdist as used by EvaluateMetrics._calculate_metrics,
// reference _dock_widget.py:1032,1038,1050): out[i][j] = ||a_i - b_j||, float64 arithmetic in numpy's
// association order, rounded once to the float32 cost matrix the LAP kernel consumes.  One 16 x 16 thread
// tile computes 64 x 64 outputs; both point blocks sit in shared memory.  HBM-bound (4 B written per pair).
__global__ void __launch_bounds__(256) pm_cdist_kernel(const double *__restrict__ a, int n1, const double *__restrict__ b,
                                                       int n2, float *__restrict__ out, int ldo) {
    __shared__ double sa[64][3], sb[64][3];
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    for (int q = threadIdx.x; q < 192; q += 256) {
        const int r = q / 3, c = q % 3;
        sa[r][c] = (i0 + r < n1) ? a[(size_t)(i0 + r) * 3 + c] : 0.0;
        sb[r][c] = (j0 + r < n2) ? b[(size_t)(j0 + r) * 3 + c] : 0.0;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r;
        if (i >= n1) continue;
        const double ax = sa[ty * 4 + r][0], ay = sa[ty * 4 + r][1], az = sa[ty * 4 + r][2];
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int jj = tx * 4 + c;
            const double d0 = ax - sb[jj][0], d1 = ay - sb[jj][1], d2 = az - sb[jj][2];
            v[c] = (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));
        }
        float *row = out + (size_t)i * ldo + j0 + tx * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (j0 + tx * 4 + c < n2) row[c] = v[c];
    }
}

extern "C" int pm_cdist(const double *a, int n1, const double *b, int n2, float *out, int ldo, void *stream) {
    PM_REQUIRE(a && b && out, "null pointer");
    PM_REQUIRE(n1 >= 1 && n2 >= 1 && ldo >= n2, "need n1, n2 >= 1 and ldo >= n2");
    dim3 grid((n2 + 63) / 64, (n1 + 63) / 64);
    pm_cdist_kernel<<<grid, 256, 0, pm_stream(stream)>>>(a, n1, b, n2, out, ldo);
    PM_LAUNCH_CHECK();
    return PM_OK;
}
