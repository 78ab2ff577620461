"""Golden vectors for transform='Similar'  from the reference, in the build container only.

TEST INFRASTRUCTURE.  `get_similar_transform` (find_transform.py:21-99) takes `q = D[0]` (:66): the first ROW of
numpy's eigenvector matrix, where Horn's method needs the eigenvector (COLUMN) of the largest eigenvalue.  Two sets
of vectors are dumped:
  *_shipped   the UNMODIFIED reference (through oracle/ref_shim.py) — what this container's numpy/LAPACK returns;
  *_fixed     the reference's own source with that one token changed (`D[0]` -> `D[:, 0]`), exec'd from
              /root/reference at generation time (nothing of it is copied into the repo) — Horn as published.
do_ransac (shape_context.py:103-139) and perform_icp (perform_icp.py:7-26) are run with transform='Similar' against
both variants by swapping the module-level `get_similar_transform` they call.

    python oracle/make_golden_similar.py        ->  tests/golden/similar.npz
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

REF = ref_shim.load()
import platymatch.estimate_transform.find_transform as ft  # noqa: E402
import platymatch.estimate_transform.perform_icp as icp_mod  # noqa: E402
import platymatch.estimate_transform.shape_context as sc_mod  # noqa: E402
from make_golden import A_GT_TEST, load_asset  # noqa: E402

src = open(os.path.join(ref_shim.REFERENCE_ROOT, "platymatch", "estimate_transform", "find_transform.py")).read()
assert src.count("q = D[0]") == 1
ns = {}
exec(compile(src.replace("q = D[0]", "q = D[:, 0]"), "find_transform_fixed", "exec"), ns)
similar_fixed = ns["get_similar_transform"]
similar_shipped = ft.get_similar_transform


def with_variant(fn, call):
    old = (sc_mod.get_similar_transform, icp_mod.get_similar_transform)
    sc_mod.get_similar_transform = icp_mod.get_similar_transform = fn
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return call()
    finally:
        sc_mod.get_similar_transform, icp_mod.get_similar_transform = old


def main():
    out = {}
    rng = np.random.default_rng(7)
    moving = load_asset("02-insitu.csv")
    # a true similarity as ground truth: rotation part of the reference tests' matrix x 1.25
    u, _, vt = np.linalg.svd(A_GT_TEST[:3, :3])
    a_gt = np.eye(4)
    a_gt[:3, :3] = 1.25 * (u @ vt)
    a_gt[:3, 3] = A_GT_TEST[:3, 3]
    fixed = (a_gt @ np.vstack([moving, np.ones((1, moving.shape[1]))]))[:3]
    out.update(moving=moving, fixed=fixed, A_gt=a_gt)
    # direct fits: exact similarity (K = 4, 10, all) and noisy pairs
    for tag, k, noise in (("k4", 4, 0.0), ("k10", 10, 0.0), ("all", moving.shape[1], 0.0), ("k12n", 12, 2.0), ("alln", moving.shape[1], 3.0)):
        sel = rng.permutation(moving.shape[1])[:k]
        m, f = moving[:, sel], fixed[:, sel] + rng.normal(0, noise, size=(3, k)) if noise else fixed[:, sel]
        out["fit_m_" + tag], out["fit_f_" + tag] = m, f
        out["fit_shipped_" + tag] = similar_shipped(m, f)
        out["fit_fixed_" + tag] = similar_fixed(m, f)
    # do_ransac, transform='Similar': correspondences = identity with 30 % of them scrambled, pinned RNG
    k = moving.shape[1]
    f_corr = fixed.copy()
    bad = rng.permutation(k)[: int(0.3 * k)]
    f_corr[:, bad] = fixed[:, rng.permutation(bad)]
    out["ransac_f"] = f_corr
    trials, seed = 300, 11
    out.update(ransac_trials=trials, ransac_seed=seed)
    for tag, fn in (("shipped", similar_shipped), ("fixed", similar_fixed)):
        np.random.seed(seed)
        A, inl = with_variant(fn, lambda: REF.do_ransac(moving, f_corr, 4, trials, 16, "Similar"))
        out["ransac_A_" + tag], out["ransac_inliers_" + tag] = A, inl
    # perform_icp, transform='Similar', from a perturbed start
    pert = np.eye(4)
    pert[:3, :3] = 1.03 * np.array([[np.cos(0.05), -np.sin(0.05), 0], [np.sin(0.05), np.cos(0.05), 0], [0, 0, 1]])
    pert[:3, 3] = [3.0, -2.0, 4.0]
    start = (pert @ a_gt @ np.vstack([moving, np.ones((1, k))]))[:3]
    out["icp_start"] = start
    for tag, fn in (("shipped", similar_shipped), ("fixed", similar_fixed)):
        out["icp_A_" + tag] = with_variant(fn, lambda: REF.perform_icp(start, fixed, 20, "Similar"))
    path = os.path.join(ROOT, "tests", "golden", "similar.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    print("ransac inliers shipped / fixed:", out["ransac_inliers_shipped"], out["ransac_inliers_fixed"], "of", k)
    print("fit all, fixed vs gt:", np.abs(out["fit_fixed_all"] - a_gt).max(), " shipped vs gt:", np.abs(out["fit_shipped_all"] - a_gt).max())
    print("icp fixed @ pert^-1... residual to gt:", np.abs(out["icp_A_fixed"] @ pert - np.eye(4)).max(),
          " shipped:", np.abs(out["icp_A_shipped"] @ pert - np.eye(4)).max())


if __name__ == "__main__":
    main()
