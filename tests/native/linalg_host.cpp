// Host build of platymatch_b200/csrc/pm_linalg.cuh (the kernels' 4x4 linear algebra) for the CPU tests:
// tests/test_linalg_host.py pins it against numpy.linalg.pinv / Horn's method without a GPU.
#include "../../platymatch_b200/csrc/pm_linalg.cuh"

extern "C" {
void t_affine_from_pairs(const double *moving, const double *fixed, int k, int use_shift, double rel_tol, double *A) {
    double shift[3] = {0, 0, 0};
    if (use_shift) { shift[0] = moving[0]; shift[1] = moving[1]; shift[2] = moving[2]; }
    double M[16], FM[12];
    pm_normal_eq_accumulate(moving, fixed, k, shift, M, FM);
    pm_affine_from_normal_eq(M, FM, shift, rel_tol, A);
}
void t_jacobi_sym4(const double *A_in, double *V, double *w) {
    double A[16];
    for (int i = 0; i < 16; ++i) A[i] = A_in[i];
    pm_jacobi_sym4(A, V, w);
}
void t_similar_from_pairs(const double *moving, const double *fixed, int k, double *A) {
    pm_similar_from_pairs(moving, fixed, k, A);
}
}
