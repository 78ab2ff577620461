/*
 * pm_oracle.c — CPU restatement (plain C, float64) of the heavy loops of PlatyMatch's
 * estimate_transform path.  TEST INFRASTRUCTURE ONLY: nothing under platymatch_b200/ links,
 * loads or calls this file; it exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py have something to check the CUDA path
 * against (and to time next to it).  Parity is PINNED: tests/test_oracle_golden.py checks every
 * function here against vectors dumped from the unmodified reference (oracle/make_golden.py).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 * The linear assignment solver restates the published algorithm behind
 * scipy.optimize.linear_sum_assignment (scipy 1.18.1, third-party, not vendored in the
 * reference; call sites platymatch/_dock_widget.py:604-611): D. F. Crouse, "On implementing 2D
 * rectangular assignment algorithms", IEEE TAES 52(4), 2016 — shortest augmenting paths with
 * Dijkstra over reduced costs, rows added one at a time, duals u/v updated after every path.
 *
 * Build: oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off; no -ffast-math, no FMA contraction,
 * so the float64 arithmetic order is the one written here).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PMO_NBINS 360

int pmo_version(void) { return 1; }

int pmo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void pmo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* platymatch/utils/utils.py:58-75 get_mean_distance: mean Euclidean distance over unordered
 * pairs i<j.  pts is N x 3 row-major (zyx).  Row-wise partial sums, then summed. */
double pmo_mean_distance(const double *pts, int n) {
    double total = 0.0;
    if (n < 2) return NAN;
    double *rows = (double *)calloc((size_t)n, sizeof(double));
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int j = i + 1; j < n; ++j) {
            double d0 = pts[3 * i] - pts[3 * j], d1 = pts[3 * i + 1] - pts[3 * j + 1],
                   d2 = pts[3 * i + 2] - pts[3 * j + 2];
            s += sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        }
        rows[i] = s;
    }
    for (int i = 0; i < n; ++i) total += rows[i];
    free(rows);
    return total / (0.5 * (double)n * (double)(n - 1));
}

/* numpy float64 floor division (`//`), used by platymatch/estimate_transform/shape_context.py:51-52
 * on np.float64 scalars: quotient from an exact fmod, then snapped (numpy npy_divmod semantics). */
static double pmo_floor_divide(double a, double b) {
    if (b == 0.0) return a / b;
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0) {
        if ((b < 0.0) != (mod < 0.0)) div -= 1.0;
    }
    if (div != 0.0) {
        double fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
        return fl;
    }
    return copysign(0.0, a / b);
}

/* One neighbour in local coordinates (a,b,c) -> linear bin or -1 (uncounted).
 * shape_context.py:25-35 (r, theta, phi) and :46-58 (get_bin_index); the linear index is NOT
 * clamped: theta == pi gives theta_index 6, phi rounding to 2*pi gives phi_index 12; indices
 * outside 0..359 and NaN are never counted by `index.count(i)` (:38-40). */
static int pmo_bin_of(double a, double b, double c, double mean_dist, const double *r_edges, int n_redges) {
    const double w_theta = M_PI / 6.0;        /* np.pi / n_thetabins */
    const double w_phi = 2.0 * M_PI / 12.0;   /* 2 * np.pi / n_phibins */
    double r_ = sqrt(a * a + b * b + c * c);  /* np.linalg.norm(neighbor) */
    double r = r_ / mean_dist;
    double theta = acos(c / r_);
    double phi = atan2(b, a);
    if (phi < 0.0) phi = 2.0 * M_PI + phi;
    double ti = pmo_floor_divide(theta, w_theta);
    double pi_ = pmo_floor_divide(phi, w_phi);
    int r_index = n_redges - 1;
    for (int e = 0; e < n_redges; ++e)
        if (r < r_edges[e]) { r_index = e; break; }
    double idx = (double)r_index * 72.0 + ti * 12.0 + pi_;
    if (!(idx >= 0.0 && idx < (double)PMO_NBINS)) return -1; /* NaN fails both */
    return (int)idx;
}

/* shape_context.py:144-188 get_unary, one orientation variant.
 *   pts       N x 3 row-major;  centroid[3];  x0[3] = first PCA axis (:162-165)
 *   sign_x/sign_y: frame (sign_x * x, sign_y * y, z).  Reference variants:
 *     sc  (+1,+1)   sc2 (-1,-1)   sc3 (+1,-1)   sc4 (-1,+1)     (:170-185; x2 = -x, y2 = -y)
 *   counts    N x 360 uint32 integer histogram (the reference row is counts / counts.sum(), :41)
 *   dropped   N: neighbours that landed outside bins 0..359 or were NaN
 * Local coordinates are the dot products of (q - p) with the frame, which is what
 * transform() (:61-84, T = B inv(A)) evaluates for an orthonormal frame. */
void pmo_shape_context_counts_range(const double *pts, int n, const double *centroid, const double *x0,
                                    double mean_dist, int sign_x, int sign_y, const double *r_edges,
                                    int n_redges, int i_begin, int i_end, uint32_t *counts, uint32_t *dropped);

void pmo_shape_context_counts(const double *pts, int n, const double *centroid, const double *x0,
                              double mean_dist, int sign_x, int sign_y, const double *r_edges,
                              int n_redges, uint32_t *counts, uint32_t *dropped) {
    pmo_shape_context_counts_range(pts, n, centroid, x0, mean_dist, sign_x, sign_y, r_edges, n_redges, 0, n,
                                   counts, dropped);
}

/* Same, for query nuclei [i_begin, i_end) only (all n nuclei are still the neighbours); row i is
 * written at counts + (i - i_begin) * 360.  Used for bounded-sample CPU timings. */
void pmo_shape_context_counts_range(const double *pts, int n, const double *centroid, const double *x0,
                                    double mean_dist, int sign_x, int sign_y, const double *r_edges,
                                    int n_redges, int i_begin, int i_end, uint32_t *counts, uint32_t *dropped) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = i_begin; i < i_end; ++i) {
        const double *p = pts + 3 * i;
        double z[3], x[3], y[3];
        double dz[3] = {p[0] - centroid[0], p[1] - centroid[1], p[2] - centroid[2]};
        double nz = sqrt(dz[0] * dz[0] + dz[1] * dz[1] + dz[2] * dz[2]);
        for (int k = 0; k < 3; ++k) z[k] = dz[k] / nz;                       /* :169 */
        double xs[3] = {sign_x * x0[0], sign_x * x0[1], sign_x * x0[2]};     /* :166 */
        double dot = xs[0] * z[0] + xs[1] * z[1] + xs[2] * z[2];
        for (int k = 0; k < 3; ++k) x[k] = xs[k] - z[k] * dot;               /* :170 */
        double nx = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        for (int k = 0; k < 3; ++k) x[k] /= nx;                              /* :171 */
        y[0] = z[1] * x[2] - z[2] * x[1];                                    /* get_Y :6-8 */
        y[1] = z[2] * x[0] - z[0] * x[2];
        y[2] = z[0] * x[1] - z[1] * x[0];
        double ny = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
        /* y belongs to the frame whose x is sign_x * x0's Gram-Schmidt; the reference's sc3/sc4
         * negate that y (:180-181), so the effective sign on cross(z, x) is sign_x * sign_y. */
        double sy = (double)(sign_x * sign_y);
        for (int k = 0; k < 3; ++k) y[k] = sy * (y[k] / ny);
        uint32_t *h = counts + (size_t)(i - i_begin) * PMO_NBINS;
        memset(h, 0, PMO_NBINS * sizeof(uint32_t));
        uint32_t drop = 0;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;                                            /* np.delete :168 */
            const double *q = pts + 3 * j;
            double d[3] = {q[0] - p[0], q[1] - p[1], q[2] - p[2]};
            double a = d[0] * x[0] + d[1] * x[1] + d[2] * x[2];
            double b = d[0] * y[0] + d[1] * y[1] + d[2] * y[2];
            double c = d[0] * z[0] + d[1] * z[1] + d[2] * z[2];
            int bin = pmo_bin_of(a, b, c, mean_dist, r_edges, n_redges);
            if (bin < 0) ++drop; else ++h[bin];
        }
        if (dropped) dropped[i - i_begin] = drop;
    }
}

/* shape_context.py:88-99 get_unary_distance over all pairs (loops _dock_widget.py:556-602):
 * out[i*n2+j] = 0.5 * sum_k (a_k-b_k)^2/(a_k+b_k), bins with a_k == b_k skipped. */
void pmo_chi2_matrix(const double *a, int n1, const double *b, int n2, int nbins, double *out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < n1; ++i) {
        const double *ai = a + (size_t)i * nbins;
        for (int j = 0; j < n2; ++j) {
            const double *bj = b + (size_t)j * nbins;
            double dist = 0.0;
            for (int k = 0; k < nbins; ++k) {
                if (ai[k] != bj[k]) {
                    double d = ai[k] - bj[k];
                    dist = dist + (d * d) / (ai[k] + bj[k]);
                }
            }
            out[(size_t)i * n2 + j] = 0.5 * dist;
        }
    }
}

/* Rectangular linear sum assignment (min), nr <= nc, cost row-major nr x nc float64.
 * Restates Crouse 2016 / scipy.optimize.linear_sum_assignment (call sites
 * _dock_widget.py:604-611).  col4row[nr] out; returns 0, or -1 if infeasible / bad shape.
 * steps_out (optional): number of Dijkstra column selections (work counter for DESIGN.md). */
int pmo_lsap(const double *cost, int nr, int nc, int64_t *col4row, double *u_out, double *v_out,
             int64_t *steps_out) {
    if (nr > nc || nr < 0) return -1;
    double *u = (double *)calloc((size_t)nr + 1, sizeof(double));
    double *v = (double *)calloc((size_t)nc + 1, sizeof(double));
    double *sp = (double *)malloc(((size_t)nc + 1) * sizeof(double));
    int64_t *path = (int64_t *)malloc(((size_t)nc + 1) * sizeof(int64_t));
    int64_t *row4col = (int64_t *)malloc(((size_t)nc + 1) * sizeof(int64_t));
    int64_t *remaining = (int64_t *)malloc(((size_t)nc + 1) * sizeof(int64_t));
    char *SR = (char *)malloc((size_t)nr + 1), *SC = (char *)malloc((size_t)nc + 1);
    int64_t steps = 0;
    int rc = 0;
    for (int i = 0; i < nr; ++i) col4row[i] = -1;
    for (int j = 0; j < nc; ++j) row4col[j] = -1;
    for (int cur = 0; cur < nr && rc == 0; ++cur) {
        double min_val = 0.0;
        int64_t i = cur, sink = -1;
        int64_t num_remaining = nc;
        for (int64_t it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
        memset(SR, 0, (size_t)nr);
        memset(SC, 0, (size_t)nc);
        for (int j = 0; j < nc; ++j) sp[j] = INFINITY;
        while (sink == -1) {
            int64_t index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            const double *ci = cost + (size_t)i * nc;
            for (int64_t it = 0; it < num_remaining; ++it) {
                int64_t j = remaining[it];
                double r = min_val + ci[j] - u[i] - v[j];
                if (r < sp[j]) { path[j] = i; sp[j] = r; }
                /* ties: prefer a column that ends the path (unassigned) */
                if (sp[j] < lowest || (sp[j] == lowest && row4col[j] == -1)) { lowest = sp[j]; index = it; }
            }
            ++steps;
            min_val = lowest;
            if (min_val == INFINITY) { rc = -1; break; }
            int64_t j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (rc) break;
        u[cur] += min_val;
        for (int r = 0; r < nr; ++r)
            if (SR[r] && r != cur) u[r] += min_val - sp[col4row[r]];
        for (int j = 0; j < nc; ++j)
            if (SC[j]) v[j] -= min_val - sp[j];
        int64_t j = sink;
        for (;;) {
            int64_t r = path[j];
            row4col[j] = r;
            int64_t t = col4row[r]; col4row[r] = j; j = t;
            if (r == cur) break;
        }
    }
    if (u_out) memcpy(u_out, u, (size_t)nr * sizeof(double));
    if (v_out) memcpy(v_out, v, (size_t)nc * sizeof(double));
    if (steps_out) *steps_out = steps;
    free(u); free(v); free(sp); free(path); free(row4col); free(remaining); free(SR); free(SC);
    return rc;
}

/* perform_icp.py:15-16: i2 = argmin_j ||moving_i - fixed_j|| (scipy distance_matrix =
 * sqrt(sum |d|^2), first minimum wins).  moving n1 x 3, fixed n2 x 3 row-major. */
void pmo_nearest(const double *moving, int n1, const double *fixed, int n2, int64_t *nn, double *dist) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n1; ++i) {
        const double *m = moving + 3 * i;
        double best = INFINITY;
        int64_t arg = 0;
        for (int j = 0; j < n2; ++j) {
            double d0 = fabs(fixed[3 * j] - m[0]), d1 = fabs(fixed[3 * j + 1] - m[1]),
                   d2 = fabs(fixed[3 * j + 2] - m[2]);
            double d = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
            if (d < best) { best = d; arg = j; }
        }
        nn[i] = arg;
        if (dist) dist[i] = best;
    }
}

/* shape_context.py:130-135: inliers of one hypothesis = #{k : ||fixed_k - (A moving_k)|| <= error}.
 * moving/fixed K x 3 row-major in correspondence order, A 4x4 row-major; many hypotheses. */
void pmo_ransac_score(const double *moving, const double *fixed, int k, const double *A, int trials,
                      double error, int32_t *inliers) {
#pragma omp parallel for schedule(static)
    for (int t = 0; t < trials; ++t) {
        const double *a = A + 16 * (size_t)t;
        int32_t cnt = 0;
        for (int p = 0; p < k; ++p) {
            const double *m = moving + 3 * p, *f = fixed + 3 * p;
            double e0 = f[0] - (((a[0] * m[0] + a[1] * m[1]) + a[2] * m[2]) + a[3]);
            double e1 = f[1] - (((a[4] * m[0] + a[5] * m[1]) + a[6] * m[2]) + a[7]);
            double e2 = f[2] - (((a[8] * m[0] + a[9] * m[1]) + a[10] * m[2]) + a[11]);
            double d = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
            if (d <= error) ++cnt;
        }
        inliers[t] = cnt;
    }
}
