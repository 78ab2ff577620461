"""CPU pinning of the kernels' 4x4 linear algebra (platymatch_b200/csrc/pm_linalg.cuh, compiled for the host by
tests/native/Makefile) against numpy: the pseudo-inverse affine fit of find_transform.py:4-17 including
rank-deficient point sets (coplanar keypoints), the Jacobi eigen-solver, and Horn's similarity fit
(find_transform.py:21-99, with the published column convention)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def L():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "native")])
    lib = ctypes.CDLL(os.path.join(HERE, "native", "liblinalg_host.so"))
    vp, i, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    lib.t_affine_from_pairs.argtypes = [vp, vp, i, i, d, vp]
    lib.t_jacobi_sym4.argtypes = [vp, vp, vp]
    lib.t_similar_from_pairs.argtypes = [vp, vp, i, vp]
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _affine(L, m, f, shift=1, tol=1e-14):
    m = np.ascontiguousarray(m.T)
    f = np.ascontiguousarray(f.T)
    A = np.empty(16)
    L.t_affine_from_pairs(_p(m), _p(f), m.shape[0], shift, tol, _p(A))
    return A.reshape(4, 4)


def _pinv_affine(m, f):
    ones = np.ones((1, m.shape[1]))
    return np.vstack([f, ones]) @ np.linalg.pinv(np.vstack([m, ones]))


def test_jacobi_matches_eigh(L):
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.normal(size=(4, 4)) * 10 ** rng.uniform(-3, 6)
        a = a + a.T
        V, w = np.empty(16), np.empty(4)
        L.t_jacobi_sym4(_p(np.ascontiguousarray(a)), _p(V), _p(w))
        V = V.reshape(4, 4)
        assert np.allclose(np.sort(w), np.linalg.eigvalsh(a), rtol=1e-12, atol=1e-12 * np.abs(a).max())
        assert np.allclose(V @ np.diag(w) @ V.T, a, rtol=1e-12, atol=1e-12 * np.abs(a).max())
        assert np.allclose(V.T @ V, np.eye(4), atol=1e-13)


@pytest.mark.parametrize("k", [4, 10, 500])
def test_affine_full_rank_equals_pinv(L, k):
    rng = np.random.default_rng(k)
    m = rng.normal(size=(3, k)) * 80 + np.array([[400.0], [300.0], [250.0]])
    f = rng.normal(size=(3, k)) * 80 + 300
    ref = _pinv_affine(m, f)
    assert np.allclose(_affine(L, m, f), ref, rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("case", ["one_slice", "tilted_plane", "line", "three_points", "two_points", "single", "duplicates"])
def test_affine_rank_deficient_equals_pinv_min_norm(L, case):
    """numpy's pinv defines the fit for coplanar / collinear / too few points (minimum-norm solution); the kernels'
    normal-equation solve falls back to the same answer instead of NaN."""
    rng = np.random.default_rng(len(case))
    k = 12
    m = rng.normal(size=(3, k)) * 60 + np.array([[300.0], [200.0], [150.0]])
    if case == "one_slice":
        m[0] = 37.0                                   # keypoints picked in one z slice
    elif case == "tilted_plane":
        m[2] = 0.3 * m[0] - 0.7 * m[1] + 11.0
    elif case == "line":
        t = rng.normal(size=k) * 50
        m = np.array([[100.0], [50.0], [20.0]]) + np.outer([1.0, 2.0, -0.5], t)
    elif case == "three_points":
        m = m[:, :3]
    elif case == "two_points":
        m = m[:, :2]
    elif case == "single":
        m = m[:, :1]
    elif case == "duplicates":
        m = np.repeat(m[:, :2], 5, axis=1)
    f = rng.normal(size=m.shape) * 60 + 250
    ref = _pinv_affine(m, f)
    got = _affine(L, m, f)
    scale = np.abs(ref).max()
    assert np.all(np.isfinite(got))
    assert np.allclose(got, ref, rtol=1e-7, atol=1e-7 * scale), np.abs(got - ref).max()
    got0 = _affine(L, m, f, shift=0)                  # the same without the conditioning shift
    assert np.allclose(got0, ref, rtol=1e-6, atol=1e-6 * scale)


def _horn_numpy(moving, fixed):
    """Horn's closed form as published (find_transform.py:21-99 with q = the eigenvector COLUMN of the largest
    eigenvalue)."""
    ct, cs = fixed.mean(1, keepdims=True), moving.mean(1, keepdims=True)
    Y, P = fixed - ct, moving - cs
    S = P @ Y.T
    Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz = S.reshape(-1)
    N = np.array([[Sxx + Syy + Szz, Syz - Szy, -Sxz + Szx, Sxy - Syx],
                  [-Szy + Syz, Sxx - Szz - Syy, Sxy + Syx, Sxz + Szx],
                  [Szx - Sxz, Syx + Sxy, Syy - Szz - Sxx, Syz + Szy],
                  [-Syx + Sxy, Szx + Sxz, Szy + Syz, Szz - Syy - Sxx]])
    w, v = np.linalg.eigh(N)
    q0, q1, q2, q3 = v[:, -1]
    Qbar = np.array([[q0, -q1, -q2, -q3], [q1, q0, q3, -q2], [q2, -q3, q0, q1], [q3, q2, -q1, q0]])
    Q = np.array([[q0, -q1, -q2, -q3], [q1, q0, -q3, q2], [q2, q3, q0, -q1], [q3, -q2, q1, q0]])
    R = (Qbar.T @ Q)[1:, 1:]
    s = np.sqrt((Y * Y).sum() / (P * P).sum())
    A = np.zeros((4, 4))
    A[:3, :3] = s * R
    A[:3, 3:4] = ct - s * R @ cs
    A[3, 3] = 1
    return A


@pytest.mark.parametrize("k", [3, 4, 50])
def test_similar_recovers_known_similarity(L, k):
    rng = np.random.default_rng(k)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    s, t = 1.7, np.array([[30.0], [-20.0], [5.0]])
    m = rng.normal(size=(3, k)) * 70 + 200
    f = s * q @ m + t
    A = np.empty(16)
    L.t_similar_from_pairs(_p(np.ascontiguousarray(m.T)), _p(np.ascontiguousarray(f.T)), k, _p(A))
    A = A.reshape(4, 4)
    assert np.allclose(A[:3, :3], s * q, atol=1e-9) and np.allclose(A[:3, 3:4], t, atol=1e-7)
    noisy = f + rng.normal(size=f.shape)
    L.t_similar_from_pairs(_p(np.ascontiguousarray(m.T)), _p(np.ascontiguousarray(noisy.T)), k, _p(A))
    assert np.allclose(A.reshape(4, 4), _horn_numpy(m, noisy), rtol=1e-9, atol=1e-8)
