// pm_icp.cu — K6: ICP refinement with affine re-fit.
//
// Reference: platymatch/estimate_transform/perform_icp.py:7-26.  Per iteration:
//   cost = distance_matrix(moving^T, fixed^T); i2 = argmin(cost, 1)          (:15-16, first min wins)
//   A_est = get_affine_transform(moving, fixed[:, i2])                        (:18, all N1 pairs)
//   moving = apply_affine_transform(moving, A_est)                            (:23)
//   residual = get_error(moving, fixed[:, i2])  (printed, :24; utils.py:77-88)
//   A_icp = A_est @ A_icp                                                      (:25)
// Always exactly `iterations` iterations, no convergence test, no outlier rejection.
//
// float64: brute-force tiled nearest neighbour (the N1 x N2 distance matrix is never materialised)
// fused with the per-CTA partial sums of the normal equations; one small CTA reduces the partials in
// a fixed order, solves the 4x4 system and composes; a third kernel applies the step and emits the
// residual partials.  Three launches per iteration, no host synchronisation inside the loop.
#include "pm_common.cuh"

#define PM_ICP_PTS 16      // moving points per CTA
#define PM_ICP_SLICES 16   // column slices per moving point (threads = 16 x 16)
#define PM_ICP_TILE 512    // fixed points staged per tile
#define PM_ICP_NSUM 22     // 10 (M M^T upper) + 12 (F M^T)

// cur: current moving positions [n1][3]; shift: 3 doubles subtracted from moving coordinates when
// forming the normal equations (conditioning); partial: [gridDim.x][PM_ICP_NSUM].
__global__ void __launch_bounds__(PM_ICP_PTS * PM_ICP_SLICES)
pm_icp_nn_kernel(const double *__restrict__ cur, int n1, const double *__restrict__ fixed, int n2,
                 const double *__restrict__ shift, int32_t *__restrict__ nn, double *__restrict__ partial) {
    __shared__ __align__(16) double tile[2][PM_ICP_TILE * 3];      // double-buffered with cp.async
    __shared__ double best_d[PM_ICP_SLICES][PM_ICP_PTS];
    __shared__ int best_j[PM_ICP_SLICES][PM_ICP_PTS];
    __shared__ double sums[PM_ICP_PTS][PM_ICP_NSUM + 1];
    const int p = threadIdx.x % PM_ICP_PTS, s = threadIdx.x / PM_ICP_PTS;
    const int i = blockIdx.x * PM_ICP_PTS + p;
    const bool live = i < n1;
    double mx = 0, my = 0, mz = 0;
    if (live) { mx = cur[3 * i]; my = cur[3 * i + 1]; mz = cur[3 * i + 2]; }
    // The reference keeps the first strict minimum of sqrt(d2) in ascending j.  sqrt is monotone, so a new
    // minimum needs d2 < best d2: the hot loop compares squared distances and only on that (rare) branch
    // evaluates the square roots, which decide exactly like the reference (two different d2 can round to
    // the same distance, and then the earlier index must stay).
    double bd = INFINITY, bd2 = INFINITY;
    int bj = 0x7fffffff;
    auto stage = [&](int buf, int t0) {        // 8-byte cp.async copies of one tile of the fixed cloud
        const int tn = min(PM_ICP_TILE, n2 - t0);
        for (int q = threadIdx.x; q < tn * 3; q += blockDim.x) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&tile[buf][q]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(fixed + (size_t)t0 * 3 + q));
        }
        asm volatile("cp.async.commit_group;");
    };
    stage(0, 0);
    int buf = 0;
    for (int t0 = 0; t0 < n2; t0 += PM_ICP_TILE, buf ^= 1) {
        const int tn = min(PM_ICP_TILE, n2 - t0);
        if (t0 + PM_ICP_TILE < n2) {
            stage(buf ^ 1, t0 + PM_ICP_TILE);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        const double *tl = tile[buf];
        int j = s;
        for (; j + PM_ICP_SLICES < tn; j += 2 * PM_ICP_SLICES) {   // two candidates per trip: independent chains
            const int ja = j, jb = j + PM_ICP_SLICES;
            const double a0 = tl[3 * ja] - mx, a1 = tl[3 * ja + 1] - my, a2 = tl[3 * ja + 2] - mz;
            const double b0 = tl[3 * jb] - mx, b1 = tl[3 * jb + 1] - my, b2 = tl[3 * jb + 2] - mz;
            const double qa = __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(a2, a2));
            const double qb = __dadd_rn(__dadd_rn(__dmul_rn(b0, b0), __dmul_rn(b1, b1)), __dmul_rn(b2, b2));
            if (qa < bd2) {                               // ascending j within a slice: first min kept
                const double d = sqrt(qa);
                if (d < bd) { bd = d; bd2 = qa; bj = t0 + ja; }
            }
            if (qb < bd2) {
                const double d = sqrt(qb);
                if (d < bd) { bd = d; bd2 = qb; bj = t0 + jb; }
            }
        }
        for (; j < tn; j += PM_ICP_SLICES) {
            const double d0 = tl[3 * j] - mx, d1 = tl[3 * j + 1] - my, d2 = tl[3 * j + 2] - mz;
            const double q2 = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
            if (q2 < bd2) {
                const double d = sqrt(q2);
                if (d < bd) { bd = d; bd2 = q2; bj = t0 + j; }
            }
        }
        __syncthreads();
    }
    best_d[s][p] = bd;
    best_j[s][p] = bj;
    __syncthreads();
    if (s == 0) {
        for (int q = 1; q < PM_ICP_SLICES; ++q) {
            const double od = best_d[q][p];
            const int oj = best_j[q][p];
            if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }   // first minimum overall
        }
        double v[PM_ICP_NSUM];
#pragma unroll
        for (int q = 0; q < PM_ICP_NSUM; ++q) v[q] = 0.0;
        if (live) {
            nn[i] = bj;
            const double m[4] = {mx - shift[0], my - shift[1], mz - shift[2], 1.0};
            const double f[3] = {fixed[3 * (size_t)bj], fixed[3 * (size_t)bj + 1], fixed[3 * (size_t)bj + 2]};
            int q = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = a; b < 4; ++b) v[q++] = m[a] * m[b];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) v[q++] = f[a] * m[b];
        }
#pragma unroll
        for (int q = 0; q < PM_ICP_NSUM; ++q) sums[p][q] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < PM_ICP_NSUM) {   // fixed-order sum over the 32 points of this CTA
        double acc = 0.0;
        for (int q = 0; q < PM_ICP_PTS; ++q) acc += sums[q][threadIdx.x];
        partial[(size_t)blockIdx.x * PM_ICP_NSUM + threadIdx.x] = acc;
    }
}

// One CTA: reduce partials, solve, compose.  A_est -> a_est[16]; a_icp <- A_est @ a_icp.
__global__ void __launch_bounds__(256) pm_icp_solve_kernel(const double *__restrict__ partial, int nblocks,
                                                           const double *__restrict__ shift,
                                                           double *__restrict__ a_est, double *__restrict__ a_icp) {
    __shared__ double tot[PM_ICP_NSUM];
    __shared__ double red[8][PM_ICP_NSUM];
    const int q = threadIdx.x % 32, g = threadIdx.x / 32;   // 8 groups x 32 lanes (22 used)
    if (q < PM_ICP_NSUM) {
        double acc = 0.0;
        for (int b = g; b < nblocks; b += 8) acc += partial[(size_t)b * PM_ICP_NSUM + q];
        red[g][q] = acc;
    }
    __syncthreads();
    if (threadIdx.x < PM_ICP_NSUM) {
        double acc = 0.0;
        for (int k = 0; k < 8; ++k) acc += red[k][threadIdx.x];
        tot[threadIdx.x] = acc;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double M[16], X[12], A[16];
    int k = 0;
    for (int a = 0; a < 4; ++a)
        for (int b = a; b < 4; ++b) { M[a * 4 + b] = tot[k]; M[b * 4 + a] = tot[k]; ++k; }
    const bool ok = pm_solve_right_4x4(M, tot + 10, 3, X, 1e-14);
    for (int r = 0; r < 3; ++r) {
        if (ok) {
            A[r * 4 + 0] = X[r * 4 + 0]; A[r * 4 + 1] = X[r * 4 + 1]; A[r * 4 + 2] = X[r * 4 + 2];
            A[r * 4 + 3] = X[r * 4 + 3] - (X[r * 4 + 0] * shift[0] + X[r * 4 + 1] * shift[1] + X[r * 4 + 2] * shift[2]);
        } else {
            for (int c = 0; c < 4; ++c) A[r * 4 + c] = nan("");
        }
    }
    A[12] = 0.0; A[13] = 0.0; A[14] = 0.0; A[15] = 1.0;
    double C[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double sacc = 0.0;
            for (int kk = 0; kk < 4; ++kk) sacc += A[r * 4 + kk] * a_icp[kk * 4 + c];
            C[r * 4 + c] = sacc;
        }
    for (int e = 0; e < 16; ++e) { a_est[e] = A[e]; a_icp[e] = C[e]; }
}

// cur <- A_est cur; residual partial per CTA (sum of ||cur' - fixed[nn]||).
__global__ void __launch_bounds__(256) pm_icp_apply_kernel(double *__restrict__ cur, int n1,
                                                           const double *__restrict__ fixed,
                                                           const int32_t *__restrict__ nn,
                                                           const double *__restrict__ a_est,
                                                           double *__restrict__ res_partial) {
    __shared__ double red[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double r = 0.0;
    if (i < n1) {
        const double x = cur[3 * i], y = cur[3 * i + 1], z = cur[3 * i + 2];
        const double nx = ((a_est[0] * x + a_est[1] * y) + a_est[2] * z) + a_est[3];
        const double ny = ((a_est[4] * x + a_est[5] * y) + a_est[6] * z) + a_est[7];
        const double nz = ((a_est[8] * x + a_est[9] * y) + a_est[10] * z) + a_est[11];
        cur[3 * i] = nx; cur[3 * i + 1] = ny; cur[3 * i + 2] = nz;
        const int j = nn[i];
        const double e0 = nx - fixed[3 * (size_t)j], e1 = ny - fixed[3 * (size_t)j + 1], e2 = nz - fixed[3 * (size_t)j + 2];
        r = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
    }
    r = pm_block_sum(r, red);
    if (threadIdx.x == 0) res_partial[blockIdx.x] = r;
}

__global__ void __launch_bounds__(256) pm_icp_residual_kernel(const double *__restrict__ res_partial, int nblocks,
                                                              int n1, double *__restrict__ residuals) {
    __shared__ double red[32];
    const double *p = res_partial + (size_t)blockIdx.x * nblocks;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += p[b];
    s = pm_block_sum(s, red);
    if (threadIdx.x == 0) residuals[blockIdx.x] = s / (double)n1;
}

__global__ void pm_icp_init_kernel(const double *__restrict__ moving, double *__restrict__ a_icp,
                                   double *__restrict__ shift) {
    if (threadIdx.x < 16) a_icp[threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.0 : 0.0;
    if (threadIdx.x < 3) shift[threadIdx.x] = moving[threadIdx.x];
}

static inline size_t pm_align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t pm_icp_workspace_bytes(int n1) {
    if (n1 < 1) return 0;
    const size_t nb_nn = (size_t)(n1 + PM_ICP_PTS - 1) / PM_ICP_PTS;
    const size_t nb_ap = (size_t)(n1 + 255) / 256;
    return pm_align256((size_t)n1 * 3 * sizeof(double))          // cur
         + pm_align256((size_t)n1 * sizeof(int32_t))             // nn
         + pm_align256(nb_nn * PM_ICP_NSUM * sizeof(double))     // partial
         + pm_align256(64 * sizeof(double))                      // a_est, a_icp, shift
         + pm_align256(nb_ap * sizeof(double) * 1024);           // residual partials (<= 1024 iterations)
}

extern "C" int pm_icp_affine(const double *moving, int n1, const double *fixed, int n2, int iterations,
                             double *A_icp, double *residuals, int32_t *nn_out, void *workspace,
                             size_t workspace_bytes, void *stream) {
    PM_REQUIRE(moving && fixed && A_icp && workspace, "null pointer");
    PM_REQUIRE(n1 >= 4 && n2 >= 1, "need n1 >= 4 and n2 >= 1");
    PM_REQUIRE(iterations >= 0 && iterations <= 1024, "iterations must be 0..1024");
    if (workspace_bytes < pm_icp_workspace_bytes(n1)) {
        pm_set_error("pm_icp_affine: workspace too small");
        return PM_ERR_WORKSPACE;
    }
    cudaStream_t s = pm_stream(stream);
    const int nb_nn = (n1 + PM_ICP_PTS - 1) / PM_ICP_PTS, nb_ap = (n1 + 255) / 256;
    char *w = (char *)workspace;
    double *cur = (double *)w; w += pm_align256((size_t)n1 * 3 * sizeof(double));
    int32_t *nn = (int32_t *)w; w += pm_align256((size_t)n1 * sizeof(int32_t));
    double *partial = (double *)w; w += pm_align256((size_t)nb_nn * PM_ICP_NSUM * sizeof(double));
    double *small = (double *)w; w += pm_align256(64 * sizeof(double));
    double *res_partial = (double *)w;
    double *a_est = small, *a_icp = small + 16, *shift = small + 32;
    PM_CUDA_TRY(cudaMemcpyAsync(cur, moving, (size_t)n1 * 3 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    pm_icp_init_kernel<<<1, 32, 0, s>>>(moving, a_icp, shift);
    PM_LAUNCH_CHECK();
    for (int it = 0; it < iterations; ++it) {
        pm_icp_nn_kernel<<<nb_nn, PM_ICP_PTS * PM_ICP_SLICES, 0, s>>>(cur, n1, fixed, n2, shift, nn, partial);
        pm_icp_solve_kernel<<<1, 256, 0, s>>>(partial, nb_nn, shift, a_est, a_icp);
        pm_icp_apply_kernel<<<nb_ap, 256, 0, s>>>(cur, n1, fixed, nn, a_est, res_partial + (size_t)it * nb_ap);
    }
    PM_LAUNCH_CHECK_N(3 * iterations);
    if (residuals && iterations > 0) {
        pm_icp_residual_kernel<<<iterations, 256, 0, s>>>(res_partial, nb_ap, n1, residuals);
        PM_LAUNCH_CHECK();
    }
    PM_CUDA_TRY(cudaMemcpyAsync(A_icp, a_icp, 16 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (nn_out && iterations > 0)
        PM_CUDA_TRY(cudaMemcpyAsync(nn_out, nn, (size_t)n1 * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    return PM_OK;
}
