"""CPU oracle for the estimate_transform path — numpy for the small linear algebra, C
(oracle/pm_oracle.c via ctypes) for the O(N^2) loops.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  platymatch_b200/ never imports it.
Parity is PINNED: tests/test_oracle_golden.py checks these functions against vectors dumped
from the unmodified reference by oracle/make_golden.py (tests/golden/*.npz).

Function names and signatures mirror the reference (paths relative to /root/reference):
  platymatch/utils/utils.py:48-88, platymatch/estimate_transform/{shape_context,find_transform,
  apply_transform,perform_icp}.py, and the pipeline body of platymatch/_dock_widget.py:526-721.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NBINS = 360

_c_double_p = ctypes.POINTER(ctypes.c_double)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpm_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.pmo_mean_distance.restype = ctypes.c_double
        L.pmo_mean_distance.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.pmo_shape_context_counts.restype = None
        L.pmo_shape_context_counts.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.pmo_shape_context_counts_range.restype = None
        L.pmo_shape_context_counts_range.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                     ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                     ctypes.c_void_p]
        L.pmo_chi2_matrix.restype = None
        L.pmo_chi2_matrix.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p]
        L.pmo_lsap.restype = ctypes.c_int
        L.pmo_lsap.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_void_p, ctypes.c_void_p]
        L.pmo_nearest.restype = None
        L.pmo_nearest.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p]
        L.pmo_ransac_score.restype = None
        L.pmo_ransac_score.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
        L.pmo_num_threads.restype = ctypes.c_int
        L.pmo_set_num_threads.argtypes = [ctypes.c_int]
        _LIB = L
    return _LIB


def num_threads():
    return lib().pmo_num_threads()


def set_num_threads(n):
    lib().pmo_set_num_threads(int(n))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _n_by_3(detections, transposed):
    """Reference convention (shape_context.py:151-157, utils.py:64-70): 3/4 x N unless transposed."""
    d = np.asarray(detections, dtype=np.float64)
    if not transposed:
        d = d.transpose()
    return np.ascontiguousarray(d[:, :3])


# --------------------------------------------------------------------------- utils.py
def get_centroid(detections, transposed=True):
    """utils.py:48-56."""
    d = np.asarray(detections, dtype=np.float64)
    if transposed:
        return np.mean(d[:, :3], 0, keepdims=True)
    return np.mean(d[:3, :], 1, keepdims=True)


def get_mean_distance(detections, transposed=True):
    """utils.py:58-75 (mean over unordered pairs)."""
    p = _n_by_3(detections, transposed)
    return float(lib().pmo_mean_distance(_ptr(p), p.shape[0]))


def get_error(moving_landmarks, fixed_landmarks):
    """utils.py:77-88."""
    if moving_landmarks is not None or fixed_landmarks is not None:
        return float(np.mean(np.linalg.norm(moving_landmarks - fixed_landmarks, axis=0)))
    return None


# --------------------------------------------------------------------------- shape_context.py
def r_edges(r_inner=1 / 8, r_outer=2, n_rbins=5):
    """shape_context.py:24 — the exact doubles numpy produces (not exact powers of two)."""
    return np.logspace(np.log10(r_inner), np.log10(r_outer), n_rbins)


def pca_first_axis(points_n3):
    """shape_context.py:162-165: sklearn PCA(n_components=3).fit(X).components_[0].

    Restates scikit-learn's `covariance_eigh` full solver (chosen for n_samples >= 10 *
    n_features, 3 features): C = (X^T X - n mu mu^T) / (n - 1); eigh; descending order;
    svd_flip(u_based_decision=False): each component's largest-|.| entry made positive.
    """
    x = np.asarray(points_n3, dtype=np.float64)
    n = x.shape[0]
    mu = x.mean(axis=0)
    c = x.T @ x
    c -= n * np.outer(mu, mu)
    c /= n - 1
    w, v = np.linalg.eigh(c)
    comp = v[:, ::-1].T  # rows = components, descending eigenvalue
    axis = comp[0].copy()
    if axis[np.argmax(np.abs(axis))] < 0:
        axis = -axis
    return axis


VARIANT_SIGNS = {1: (1, 1), 2: (-1, -1), 3: (1, -1), 4: (-1, 1)}  # sc, sc2, sc3, sc4 (shape_context.py:170-185)


def shape_context_counts(points_n3, centroid, mean_distance, x0, variant=1, query_range=None):
    """Integer histogram (N,360) uint32 + dropped-neighbour count per nucleus, one orientation.
    query_range=(i0, i1) restricts the query nuclei (rows) for bounded-sample timings."""
    p = np.ascontiguousarray(points_n3, dtype=np.float64)
    n = p.shape[0]
    if query_range is not None:
        i0, i1 = query_range
        c = np.ascontiguousarray(np.asarray(centroid, dtype=np.float64).reshape(-1)[:3])
        x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(3))
        e = np.ascontiguousarray(r_edges())
        counts = np.zeros((i1 - i0, NBINS), dtype=np.uint32)
        dropped = np.zeros(i1 - i0, dtype=np.uint32)
        sx, sy = VARIANT_SIGNS[variant]
        lib().pmo_shape_context_counts_range(_ptr(p), n, _ptr(c), _ptr(x0), float(mean_distance), sx, sy, _ptr(e),
                                             len(e), i0, i1, _ptr(counts), _ptr(dropped))
        return counts, dropped
    c = np.ascontiguousarray(np.asarray(centroid, dtype=np.float64).reshape(-1)[:3])
    x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(3))
    e = np.ascontiguousarray(r_edges())
    counts = np.zeros((n, NBINS), dtype=np.uint32)
    dropped = np.zeros(n, dtype=np.uint32)
    sx, sy = VARIANT_SIGNS[variant]
    lib().pmo_shape_context_counts(_ptr(p), n, _ptr(c), _ptr(x0), float(mean_distance), sx, sy, _ptr(e), len(e),
                                   _ptr(counts), _ptr(dropped))
    return counts, dropped


def normalise_counts(counts):
    """shape_context.py:41: sc / sc.sum() on the float64 count vector."""
    c = counts.astype(np.float64)
    return c / c.sum(axis=1, keepdims=True)


def get_unary(centroid, mean_distance, detections, type, transposed=False):
    """shape_context.py:144-188: (sc, sc2, sc3, sc4), each (N,360) float64; sc3/sc4 empty for 'moving'."""
    p = _n_by_3(detections, transposed)
    x0 = pca_first_axis(p)
    out = []
    for variant in (1, 2, 3, 4):
        if variant > 2 and type != "fixed":
            out.append(np.array([]))
            continue
        counts, _ = shape_context_counts(p, centroid, mean_distance, x0, variant)
        out.append(normalise_counts(counts))
    return tuple(out)


def get_unary_distance(sc1, sc2):
    """shape_context.py:88-99 for one pair (scalar API)."""
    a = np.ascontiguousarray(sc1, dtype=np.float64).reshape(1, -1)
    b = np.ascontiguousarray(sc2, dtype=np.float64).reshape(1, -1)
    return float(unary_distance_matrix(a, b)[0, 0])


def unary_distance_matrix(sc_a, sc_b):
    """The double loops of _dock_widget.py:556-602: U[i,j] = get_unary_distance(sc_a[i], sc_b[j])."""
    a = np.ascontiguousarray(sc_a, dtype=np.float64)
    b = np.ascontiguousarray(sc_b, dtype=np.float64)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
    lib().pmo_chi2_matrix(_ptr(a), a.shape[0], _ptr(b), b.shape[0], a.shape[1], _ptr(out))
    return out


# --------------------------------------------------------------------------- LAP
def linear_sum_assignment(cost, return_stats=False):
    """scipy.optimize.linear_sum_assignment (call sites _dock_widget.py:604-611): rows ascending,
    min(nr, nc) pairs; wide matrices solved directly, tall ones through the transpose."""
    c = np.asarray(cost, dtype=np.float64)
    nr, nc = c.shape
    transposed = nr > nc
    if transposed:
        c = c.T
        nr, nc = nc, nr
    c = np.ascontiguousarray(c)
    col4row = np.empty(nr, dtype=np.int64)
    u = np.empty(nr)
    v = np.empty(nc)
    steps = ctypes.c_int64(0)
    rc = lib().pmo_lsap(_ptr(c), nr, nc, _ptr(col4row), _ptr(u), _ptr(v), ctypes.byref(steps))
    if rc != 0:
        raise ValueError("cost matrix is infeasible")
    rows = np.arange(nr, dtype=np.int64)
    if transposed:
        order = np.argsort(col4row)
        rows, col4row = col4row[order], rows[order]
    if return_stats:
        return rows, col4row, dict(u=u, v=v, steps=steps.value)
    return rows, col4row


# --------------------------------------------------------------------------- find_transform.py / apply_transform.py
def get_affine_transform(moving, fixed, with_ones=False):
    """find_transform.py:4-17: fixed_h @ pinv(moving_h)."""
    moving = np.asarray(moving, dtype=np.float64)
    fixed = np.asarray(fixed, dtype=np.float64)
    if not with_ones:
        ones = np.ones((1, moving.shape[1]))
        moving = np.vstack((moving, ones))
        fixed = np.vstack((fixed, ones))
    return np.matmul(fixed, np.linalg.pinv(moving))


def get_similar_transform(moving, fixed, as_shipped=False):
    """find_transform.py:21-99, line by line.  as_shipped=True keeps `q = D[0]` (:66, the first ROW of numpy's
    eigenvector matrix: LAPACK-dependent signs, not a rotation); the default is Horn's method as published
    (`q = D[:, 0]`, the eigenvector of the largest eigenvalue), which is what the CUDA path implements."""
    moving = np.asarray(moving, dtype=np.float64)
    fixed = np.asarray(fixed, dtype=np.float64)
    com_target = np.mean(fixed, 1, keepdims=True)                       # :27
    com_source = np.mean(moving, 1, keepdims=True)                      # :28
    Yprime = fixed[:3, :] - com_target[:3, :]                           # :31
    Pprime = moving[:3, :] - com_source[:3, :]                          # :32
    Px, Py, Pz = Pprime[0, :], Pprime[1, :], Pprime[2, :]
    Yx, Yy, Yz = Yprime[0, :], Yprime[1, :], Yprime[2, :]
    Sxx, Sxy, Sxz = np.sum(Yx * Px), np.sum(Px * Yy), np.sum(Px * Yz)   # :44-46
    Syx, Syy, Syz = np.sum(Py * Yx), np.sum(Py * Yy), np.sum(Py * Yz)   # :48-50
    Szx, Szy, Szz = np.sum(Pz * Yx), np.sum(Pz * Yy), np.sum(Pz * Yz)   # :52-54
    Nmatrix = [[Sxx + Syy + Szz, Syz - Szy, -Sxz + Szx, Sxy - Syx],     # :56-59
               [-Szy + Syz, Sxx - Szz - Syy, Sxy + Syx, Sxz + Szx],
               [Szx - Sxz, Syx + Sxy, Syy - Szz - Sxx, Syz + Szy],
               [-Syx + Sxy, Szx + Sxz, Szy + Syz, Szz - Syy - Sxx]]
    V, D = np.linalg.eig(Nmatrix)                                       # :61
    idx = V.argsort()[::-1]                                             # :63
    D = D[:, idx]                                                       # :65
    q = D[0] if as_shipped else D[:, 0]                                 # :66
    q0, q1, q2, q3 = q[0], q[1], q[2], q[3]
    Qbar = [[q0, -q1, -q2, -q3], [q1, q0, q3, -q2], [q2, -q3, q0, q1], [q3, q2, -q1, q0]]     # :72-75
    Q = [[q0, -q1, -q2, -q3], [q1, q0, -q3, q2], [q2, q3, q0, -q1], [q3, -q2, q1, q0]]        # :77-80
    R = np.matmul(np.transpose(Qbar), Q)[1:, 1:]                        # :82-83
    s = np.sqrt(np.sum(Yprime * Yprime) / np.sum(Pprime * Pprime))      # :86-94
    t = com_target[:3, :] - s * np.matmul(R, com_source[:3, :])         # :95
    A = np.zeros((4, 4))
    A[:3, :3] = s * R
    A[:3, 3:4] = t
    A[3, 3] = 1
    return A


def fit_transform(moving, fixed, transform="Affine", as_shipped=False):
    if transform == "Affine":
        return get_affine_transform(moving, fixed)
    if transform == "Similar":
        return get_similar_transform(moving, fixed, as_shipped)
    raise ValueError(transform)


def apply_affine_transform(moving, affine_transform_matrix):
    """apply_transform.py:3-17."""
    moving = np.asarray(moving, dtype=np.float64)[:3, :]
    moving = np.vstack((moving, np.ones((1, moving.shape[1]))))
    return np.matmul(affine_transform_matrix, moving)[:3, :]


# --------------------------------------------------------------------------- RANSAC
def ransac_sample_indices(k, min_samples, trials, seed):
    """The index stream do_ransac draws (shape_context.py:122) when the global numpy RNG was
    seeded with `seed`: RandomState(seed).choice(k, min_samples, replace=False) per trial."""
    rs = np.random.RandomState(seed)
    return np.stack([rs.choice(k, min_samples, replace=False) for _ in range(trials)]).astype(np.int32)


def do_ransac(moving_all, fixed_all, min_samples=4, trials=500, error=5, transform="Affine", sample_indices=None,
              seed=None, return_all=False, as_shipped=False):
    """shape_context.py:103-139 with the sample-index stream made explicit (first strictly-better wins)."""
    moving_all = np.asarray(moving_all, dtype=np.float64)[:3, :]
    fixed_all = np.asarray(fixed_all, dtype=np.float64)[:3, :]
    k = fixed_all.shape[1]
    if sample_indices is None:
        sample_indices = ransac_sample_indices(k, min_samples, trials, seed)
    trials = sample_indices.shape[0]
    mats = np.empty((trials, 4, 4))
    for t in range(trials):
        idx = sample_indices[t]
        mats[t] = fit_transform(moving_all[:, idx], fixed_all[:, idx], transform, as_shipped)
    m = np.ascontiguousarray(moving_all.T)
    f = np.ascontiguousarray(fixed_all.T)
    inl = np.zeros(trials, dtype=np.int32)
    lib().pmo_ransac_score(_ptr(m), _ptr(f), k, _ptr(np.ascontiguousarray(mats)), trials, float(error), _ptr(inl))
    a_best, inliers_best = np.ones((4, 4)), 0
    if trials and inl.max() > 0:
        t = int(np.argmax(inl))  # first maximum == first strictly-better
        a_best, inliers_best = mats[t], int(inl[t])
    if return_all:
        return a_best, inliers_best, inl, mats
    return a_best, inliers_best


# --------------------------------------------------------------------------- ICP
def nearest(moving, fixed):
    """perform_icp.py:15-16 for 3xN inputs: index of the nearest fixed point per moving point."""
    m = np.ascontiguousarray(np.asarray(moving, dtype=np.float64)[:3].T)
    f = np.ascontiguousarray(np.asarray(fixed, dtype=np.float64)[:3].T)
    nn = np.empty(m.shape[0], dtype=np.int64)
    dist = np.empty(m.shape[0])
    lib().pmo_nearest(_ptr(m), m.shape[0], _ptr(f), f.shape[0], _ptr(nn), _ptr(dist))
    return nn, dist


def perform_icp(moving, fixed, icp_iterations=50, transform="Affine", return_residuals=False, as_shipped=False):
    """perform_icp.py:7-26."""
    moving = np.asarray(moving, dtype=np.float64)[:3, :]
    fixed = np.asarray(fixed, dtype=np.float64)[:3, :]
    a_icp = np.identity(4)
    residuals = []
    for _ in range(icp_iterations):
        i2, _d = nearest(moving, fixed)
        a_est = fit_transform(moving, fixed[:, i2], transform, as_shipped)
        moving = apply_affine_transform(moving, a_est)
        residuals.append(get_error(moving, fixed[:, i2]))
        a_icp = np.matmul(a_est, a_icp)
    if return_residuals:
        return a_icp, np.array(residuals)
    return a_icp


# --------------------------------------------------------------------------- pipeline (_dock_widget.py:526-721)
HYPOTHESES = [(1, 1), (1, 2), (1, 3), (1, 4), (2, 1), (2, 2), (2, 3), (2, 4)]  # U11..U14, U21..U24


def estimate_transform_unsupervised(moving, fixed, ransac_samples=4, ransac_trials=8000, ransac_error=16,
                                    icp_iterations=50, seed=0, hypotheses=None, sample_indices=None,
                                    transform="Affine"):
    moving = np.asarray(moving, dtype=np.float64)
    fixed = np.asarray(fixed, dtype=np.float64)
    mc, fc = get_centroid(moving, False), get_centroid(fixed, False)
    md, fd = get_mean_distance(moving, False), get_mean_distance(fixed, False)
    um = get_unary(mc, md, moving, "moving")
    uf = get_unary(fc, fd, fixed, "fixed")
    hyps = HYPOTHESES if hypotheses is None else hypotheses
    res = dict(inliers=[], ransac_A=[], assignments=[], lap_cost=[])
    rs = np.random.RandomState(seed)
    for q, (a, b) in enumerate(hyps):
        U = unary_distance_matrix(um[a - 1], uf[b - 1])
        r, c = linear_sum_assignment(U)
        if sample_indices is not None:
            idx = sample_indices[q]
        else:
            idx = np.stack([rs.choice(len(r), ransac_samples, replace=False) for _ in range(ransac_trials)])
        A, inl = do_ransac(moving[:, r], fixed[:, c], ransac_samples, ransac_trials, ransac_error, transform,
                           sample_indices=idx)
        res["inliers"].append(inl)
        res["ransac_A"].append(A)
        res["assignments"].append((r, c))
        res["lap_cost"].append(float(U[r, c].sum()))
    best = int(np.argmax(res["inliers"]))
    a_sc = res["ransac_A"][best]
    a_icp, resid = perform_icp(apply_affine_transform(moving, a_sc), fixed, icp_iterations, transform,
                               return_residuals=True)
    res.update(best=best, transform_sc=a_sc, transform_icp=a_icp, transform=a_icp @ a_sc, icp_residuals=resid)
    return res


def estimate_transform_supervised(moving, fixed, moving_keypoints, fixed_keypoints, icp_iterations=50,
                                  transform="Affine"):
    """_dock_widget.py:707-717: LS affine (or similarity, :710-711) on the keypoints, then ICP."""
    a_sc = fit_transform(moving_keypoints, fixed_keypoints, transform)
    a_icp, resid = perform_icp(apply_affine_transform(moving, a_sc), fixed, icp_iterations, transform,
                               return_residuals=True)
    return dict(transform_sc=a_sc, transform_icp=a_icp, transform=a_icp @ a_sc, icp_residuals=resid)


# ---------------------------------------------------------------------------------------------
# label image -> detections (reference platymatch/_dock_widget.py:497-521, inline
# in EstimateTransform._click_run; restated line by line, the reference offers no callable for it)
def detections_from_labels(label_image, anisotropy=1.0):
    """ids = np.unique(data), 0 dropped (:497-500); per id: z, y, x = np.where(data == id),
    centroid = (np.mean(z), np.mean(y), np.mean(x)) (:505-507), size = anisotropy * len(z) (:508).
    Returns (detections 3 x N zyx, sizes N, ids N)."""
    data = np.asarray(label_image)
    ids = np.unique(data)
    ids = ids[ids != 0]
    temp, sizes = [], []
    for id_ in ids:
        z, y, x = np.where(data == id_)
        temp.append([np.mean(z), np.mean(y), np.mean(x)])
        sizes.append(float(anisotropy) * len(z))
    det = np.asarray(temp, dtype=np.float64).reshape(-1, 3).transpose()
    return det, np.asarray(sizes, dtype=np.float64), ids


def ransac_error_from_sizes(moving_nucleus_size, fixed_nucleus_size):
    """reference _dock_widget.py:613-618."""
    if len(moving_nucleus_size) == 0 or len(fixed_nucleus_size) == 0:
        return 16
    return 0.5 * (np.average(moving_nucleus_size) ** (1 / 3) + np.average(fixed_nucleus_size) ** (1 / 3))


# ---------------------------------------------------------------------------------------------
# EvaluateMetrics._calculate_metrics (reference platymatch/_dock_widget.py:1030-1080),
# restated with scipy's cdist + linear_sum_assignment exactly as the reference calls them.
def calculate_metrics(moving_keypoints, moving_keypoint_ids, moving_detections, moving_ids,
                      fixed_keypoints, fixed_keypoint_ids, fixed_detections, fixed_ids,
                      transform_matrix_1, transform_matrix_2=None):
    from scipy.optimize import linear_sum_assignment as lsa
    from scipy.spatial.distance import cdist
    t2 = np.eye(4) if transform_matrix_2 is None else transform_matrix_2
    r, c = lsa(cdist(moving_keypoints.transpose(), moving_detections.transpose()))          # :1032-1033
    moving_dictionary = {}
    for index in r:
        moving_dictionary[moving_keypoint_ids[index]] = moving_ids[c[index]]
    r, c = lsa(cdist(fixed_keypoints.transpose(), fixed_detections.transpose()))            # :1038-1039
    fixed_dictionary = {}
    for index in r:
        fixed_dictionary[fixed_keypoint_ids[index]] = fixed_ids[c[index]]
    moved = apply_affine_transform(apply_affine_transform(moving_detections, transform_matrix_1), t2)   # :1044-1046
    row_indices, col_indices = lsa(cdist(moved.transpose(), fixed_detections.transpose()))  # :1050-1051
    row_ids, col_ids = moving_ids[row_indices], fixed_ids[col_indices]
    hits = 0
    for key in moving_dictionary.keys():
        if key in fixed_dictionary.keys():
            if col_ids[np.where(row_ids == moving_dictionary[key])] == fixed_dictionary[key]:
                hits += 1
    accuracy = hits / len(fixed_dictionary.keys())                                          # :1067
    combined = np.matmul(t2, transform_matrix_1)                                            # :1028
    moved_kp = apply_affine_transform(moving_keypoints, combined)
    distance = 0
    for i in range(moved_kp.shape[1]):
        distance += np.linalg.norm([fixed_keypoints.transpose()[np.where(fixed_keypoint_ids == moving_keypoint_ids[i]), :]
                                    - moved_kp.transpose()[i, :]])
    return accuracy, distance / len(moving_dictionary.keys())                                # :1079
