// pm_linalg.cuh — the small dense linear algebra of the path (4x4, float64), host + device.
//
// Everything here replaces a numpy.linalg call of the reference (paths relative to the reference repo):
//   numpy.linalg.pinv   platymatch/estimate_transform/find_transform.py:17   (fixed_h @ pinv(moving_h))
//   numpy.linalg.eig    platymatch/estimate_transform/find_transform.py:60   (Horn's 4x4 quaternion matrix)
// The functions are `__host__ __device__` so that tests/native/linalg_host.cpp can compile them with the host
// compiler and pin them against numpy on the CPU (tests/test_linalg_host.py); the product only calls them
// from kernels.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PM_HD __host__ __device__
#define PM_COLD __noinline__          // cold paths stay out of the hot kernels' register budget
#else
#define PM_HD
#define PM_COLD
#endif

// Solve X * M = R for X (rows of R independent), i.e. X = R * inv(M), by Gauss-Jordan with partial pivoting
// on M^T.  Returns false if a pivot falls below rel_tol x (largest entry): numerically singular.
PM_HD inline bool pm_solve_right_4x4(const double M[16], const double *R, int nrows, double *X, double rel_tol) {
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

// Cyclic Jacobi eigen-decomposition of a symmetric 4x4 (row-major, overwritten).  w[c] = eigenvalue,
// V[r * 4 + c] = component r of the unit eigenvector c.  Unordered.  Converges quadratically; 12 sweeps are
// far more than float64 needs.
PM_HD inline void pm_jacobi_sym4(double A[16], double V[16], double w[4]) {
    for (int i = 0; i < 16; ++i) V[i] = 0.0;
    V[0] = V[5] = V[10] = V[15] = 1.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < 4; ++p) {
            diag += A[p * 4 + p] * A[p * 4 + p];
            for (int q = p + 1; q < 4; ++q) off += A[p * 4 + q] * A[p * 4 + q];
        }
        if (!(off > 1e-34 * diag)) break;           // (also leaves on NaN)
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                const double apq = A[p * 4 + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * 4 + q] - A[p * 4 + p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {       // A <- A J  (columns p, q)
                    const double akp = A[k * 4 + p], akq = A[k * 4 + q];
                    A[k * 4 + p] = c * akp - s * akq;
                    A[k * 4 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) {       // A <- J^T A  (rows p, q)
                    const double apk = A[p * 4 + k], aqk = A[q * 4 + k];
                    A[p * 4 + k] = c * apk - s * aqk;
                    A[q * 4 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k * 4 + p], vkq = V[k * 4 + q];
                    V[k * 4 + p] = c * vkp - s * vkq;
                    V[k * 4 + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int c = 0; c < 4; ++c) w[c] = A[c * 4 + c];
}

// Least-squares affine from the normal equations of K >= 1 point pairs, == fixed_h @ pinv(moving_h)
// (find_transform.py:14-17) INCLUDING rank-deficient point sets (coplanar / collinear / fewer than 4 points),
// where pinv returns the minimum-norm solution.
//   Msym  sum m' m'^T, 4x4 symmetric, m' = [p - shift; 1]   (shifted coordinates: conditioning)
//   FM    sum f m'^T, 3x4
//   A     4x4 row-major out, in UNSHIFTED coordinates
// Full rank: Gauss-Jordan; last row exactly [0 0 0 1] (numpy's is that up to 1e-16 noise).
// Rank deficient (a pivot below rel_tol, or eigenvalues below ~1e-12 of the largest): spectral pseudo-inverse of
// Msym for one least-squares solution A0, then the minimum-norm one of the UNSHIFTED problem: all solutions are
// A0 + z n^T with n spanning the left null space of moving_h (the plane's homogeneous normal), the shortest has
// rows orthogonal to n; the last row is e4 projected the same way (what ones @ pinv(moving_h) gives).
PM_HD PM_COLD inline void pm_affine_min_norm(const double Msym[16], const double FM[12], const double shift[3],
                                             double A[16]) {
    double S[16], V[16], w[4];
    for (int i = 0; i < 16; ++i) S[i] = Msym[i];
    pm_jacobi_sym4(S, V, w);
    double wmax = 0.0;
    for (int c = 0; c < 4; ++c) wmax = fmax(wmax, fabs(w[c]));
    if (!(wmax > 0.0) || !isfinite(wmax)) {
        for (int i = 0; i < 16; ++i) A[i] = nan("");
        return;
    }
    const double cut = 1e-12 * wmax;
    double A0[16];                                  // rows 0..2: least-squares solution; row 3: e4 (exact solution)
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            double acc = 0.0;
            for (int e = 0; e < 4; ++e) {
                if (!(w[e] > cut)) continue;
                double fq = 0.0;
                for (int k = 0; k < 4; ++k) fq += FM[r * 4 + k] * V[k * 4 + e];
                acc += fq * V[c * 4 + e] / w[e];
            }
            A0[r * 4 + c] = acc;
        }
    for (int r = 0; r < 3; ++r)                     // unshift: A0 <- A0 S,  S = [[I, -shift], [0, 1]]
        A0[r * 4 + 3] -= A0[r * 4 + 0] * shift[0] + A0[r * 4 + 1] * shift[1] + A0[r * 4 + 2] * shift[2];
    A0[12] = 0.0; A0[13] = 0.0; A0[14] = 0.0; A0[15] = 1.0;
    double nul[3][4];
    int nn = 0;
    for (int e = 0; e < 4 && nn < 3; ++e) {
        if (w[e] > cut) continue;
        double n[4] = {V[0 * 4 + e], V[1 * 4 + e], V[2 * 4 + e], V[3 * 4 + e]};
        n[3] -= n[0] * shift[0] + n[1] * shift[1] + n[2] * shift[2];       // n = S^T n'
        for (int p = 0; p < nn; ++p) {              // Gram-Schmidt against the previous null vectors
            double d = 0.0;
            for (int k = 0; k < 4; ++k) d += n[k] * nul[p][k];
            for (int k = 0; k < 4; ++k) n[k] -= d * nul[p][k];
        }
        double len = 0.0;
        for (int k = 0; k < 4; ++k) len += n[k] * n[k];
        len = sqrt(len);
        if (!(len > 0.0)) continue;
        for (int k = 0; k < 4; ++k) nul[nn][k] = n[k] / len;
        ++nn;
    }
    for (int r = 0; r < 4; ++r) {
        double row[4] = {A0[r * 4 + 0], A0[r * 4 + 1], A0[r * 4 + 2], A0[r * 4 + 3]};
        for (int p = 0; p < nn; ++p) {
            double d = 0.0;
            for (int k = 0; k < 4; ++k) d += row[k] * nul[p][k];
            for (int k = 0; k < 4; ++k) row[k] -= d * nul[p][k];
        }
        for (int k = 0; k < 4; ++k) A[r * 4 + k] = row[k];
    }
}

PM_HD inline void pm_affine_from_normal_eq(const double Msym[16], const double FM[12], const double shift[3],
                                           double rel_tol, double A[16]) {
    double X[12];
    if (pm_solve_right_4x4(Msym, FM, 3, X, rel_tol)) {
        for (int r = 0; r < 3; ++r) {
            A[r * 4 + 0] = X[r * 4 + 0]; A[r * 4 + 1] = X[r * 4 + 1]; A[r * 4 + 2] = X[r * 4 + 2];
            A[r * 4 + 3] = X[r * 4 + 3] - (X[r * 4 + 0] * shift[0] + X[r * 4 + 1] * shift[1] + X[r * 4 + 2] * shift[2]);
        }
        A[12] = 0.0; A[13] = 0.0; A[14] = 0.0; A[15] = 1.0;
        return;
    }
    pm_affine_min_norm(Msym, FM, shift, A);
}

// Accumulate the normal equations of K pairs (host-side helper for the CPU pinning test; kernels do this with
// block reductions).
PM_HD inline void pm_normal_eq_accumulate(const double *moving, const double *fixed, int k, const double shift[3],
                                          double Msym[16], double FM[12]) {
    for (int i = 0; i < 16; ++i) Msym[i] = 0.0;
    for (int i = 0; i < 12; ++i) FM[i] = 0.0;
    for (int i = 0; i < k; ++i) {
        const double m[4] = {moving[3 * i] - shift[0], moving[3 * i + 1] - shift[1], moving[3 * i + 2] - shift[2], 1.0};
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) Msym[a * 4 + b] += m[a] * m[b];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 4; ++b) FM[a * 4 + b] += fixed[3 * i + a] * m[b];
    }
}

// Similarity transform (scale, rotation, translation) of K >= 3 point pairs by Horn's closed form —
// get_similar_transform, find_transform.py:21-99, as PUBLISHED: the rotation is the unit quaternion = the
// eigenvector of the LARGEST eigenvalue of the 4x4 matrix N (:54-63).  The reference as shipped takes `q = D[0]`
// (:66), the first ROW of numpy's eigenvector matrix (first components of four different eigenvectors, each with
// the arbitrary sign LAPACK's dgeev happened to give it), which is not a rotation and not reproducible by another
// eigen-solver; see DESIGN.md "Similar".  Scale = sqrt(sum |y'|^2 / sum |p'|^2) (:88-94), t = com_y - s R com_p (:95).
//   com_p, com_y  centroids of moving / fixed (:27-28)
//   S[a * 3 + b]  sum_i p'_a y'_b over the centred coordinates (:44-54: Sxy = sum(Px * Yy))
//   spp, syy      sum |p'|^2, sum |y'|^2
PM_HD inline void pm_similar_from_moments(const double com_p[3], const double com_y[3], const double S[9], double spp,
                                          double syy, double A[16]) {
    const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5], Szx = S[6], Szy = S[7], Szz = S[8];
    double N[16] = {Sxx + Syy + Szz, Syz - Szy,       -Sxz + Szx,      Sxy - Syx,
                    -Szy + Syz,      Sxx - Szz - Syy, Sxy + Syx,       Sxz + Szx,
                    Szx - Sxz,       Syx + Sxy,       Syy - Szz - Sxx, Syz + Szy,
                    -Syx + Sxy,      Szx + Sxz,       Szy + Syz,       Szz - Syy - Sxx};
    double V[16], w[4];
    pm_jacobi_sym4(N, V, w);
    int best = 0;
    for (int c = 1; c < 4; ++c)
        if (w[c] > w[best]) best = c;
    const double q0 = V[0 * 4 + best], q1 = V[1 * 4 + best], q2 = V[2 * 4 + best], q3 = V[3 * 4 + best];
    const double Qbar[16] = {q0, -q1, -q2, -q3,  q1, q0, q3, -q2,  q2, -q3, q0, q1,  q3, q2, -q1, q0};
    const double Q[16] = {q0, -q1, -q2, -q3,  q1, q0, -q3, q2,  q2, q3, q0, -q1,  q3, -q2, q1, q0};
    double R[9];
    for (int r = 1; r < 4; ++r)
        for (int c = 1; c < 4; ++c) {               // (Qbar^T Q)[1:, 1:]  (:80-81)
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += Qbar[k * 4 + r] * Q[k * 4 + c];
            R[(r - 1) * 3 + (c - 1)] = acc;
        }
    const double s = sqrt(syy / spp);
    for (int r = 0; r < 3; ++r) {
        double rc = 0.0;
        for (int c = 0; c < 3; ++c) {
            A[r * 4 + c] = s * R[r * 3 + c];
            rc += R[r * 3 + c] * com_p[c];
        }
        A[r * 4 + 3] = com_y[r] - s * rc;
    }
    A[12] = 0.0; A[13] = 0.0; A[14] = 0.0; A[15] = 1.0;
}

// host-side helper for the CPU pinning test: moments of K pairs, centred like the reference (two passes)
PM_HD inline void pm_similar_from_pairs(const double *moving, const double *fixed, int k, double A[16]) {
    double cp[3] = {0, 0, 0}, cy[3] = {0, 0, 0};
    for (int i = 0; i < k; ++i)
        for (int a = 0; a < 3; ++a) { cp[a] += moving[3 * i + a]; cy[a] += fixed[3 * i + a]; }
    for (int a = 0; a < 3; ++a) { cp[a] /= k; cy[a] /= k; }
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, spp = 0.0, syy = 0.0;
    for (int i = 0; i < k; ++i) {
        double p[3], y[3];
        for (int a = 0; a < 3; ++a) { p[a] = moving[3 * i + a] - cp[a]; y[a] = fixed[3 * i + a] - cy[a]; }
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) S[a * 3 + b] += p[a] * y[b];
            spp += p[a] * p[a];
            syy += y[a] * y[a];
        }
    }
    pm_similar_from_moments(cp, cy, S, spp, syy, A);
}
