"""GPU parity for SURVEY §8(f) row 1 (transform='Similar', find_transform.py:21-99 inside do_ransac
shape_context.py:128-129 and perform_icp perform_icp.py:19-20) and for rank-deficient point sets in the affine fit
(find_transform.py:17: pinv's minimum-norm answer).

Similar is built as Horn's method is published (eigenvector COLUMN of the largest eigenvalue); goldens
`*_fixed` in tests/golden/similar.npz come from the reference's own source with `q = D[0]` -> `q = D[:, 0]`
(oracle/make_golden_similar.py); the as-shipped goldens are pinned on the oracle side only (tests/test_oracle_golden.py).
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

FIT_TAGS = ["k4", "k10", "all", "k12n", "alln"]


@pytest.fixture(scope="module")
def torch():
    import torch
    torch.cuda.set_device(0)
    return torch


def test_get_similar_transform_vs_goldens(O, torch):
    from platymatch_b200.estimate_transform.find_transform import get_similar_transform
    g = load_golden("similar")
    for tag in FIT_TAGS:
        m, f = g["fit_m_" + tag], g["fit_f_" + tag]
        got = get_similar_transform(m, f)
        assert np.allclose(got, g["fit_fixed_" + tag], rtol=1e-7, atol=1e-7), tag
        assert np.allclose(got, O.get_similar_transform(m, f), rtol=1e-7, atol=1e-7), tag
    assert np.abs(get_similar_transform(g["moving"], g["fixed"]) - g["A_gt"]).max() < 1e-7
    m4 = np.vstack([g["moving"], np.ones((1, g["moving"].shape[1]))])          # 4 x N input is sliced (:31-32)
    assert np.allclose(get_similar_transform(m4, g["fixed"]), g["A_gt"], atol=1e-7)


def test_ransac_similar_vs_goldens(O, torch):
    from platymatch_b200.estimate_transform.shape_context import do_ransac
    g = load_golden("similar")
    k, trials = g["moving"].shape[1], int(g["ransac_trials"])
    rs = np.random.RandomState(int(g["ransac_seed"]))
    idx = np.stack([rs.choice(k, 4, replace=False) for _ in range(trials)])
    A, inl = do_ransac(g["moving"], g["ransac_f"], 4, trials, 16, 'Similar', sample_indices=idx)
    assert inl == int(g["ransac_inliers_fixed"])
    assert np.allclose(A, g["ransac_A_fixed"], rtol=1e-7, atol=1e-7)
    # per-trial inlier counts against the oracle, 3-sample and 6-sample similarity fits as well
    from platymatch_b200 import device as D
    md, fd = D.to_device_points(g["moving"]), D.to_device_points(g["ransac_f"])
    for ms in (3, 4, 6):
        idx = O.ransac_sample_indices(k, ms, 200, seed=ms)
        _, _, inl_o, _ = O.do_ransac(g["moving"], g["ransac_f"], ms, 200, 16, "Similar", sample_indices=idx, return_all=True)
        _, _, _, per = D.ransac(md, fd, 200, 16.0, ms, torch.from_numpy(idx).cuda(), want_per_trial=True, transform='Similar')
        assert np.array_equal(per.cpu().numpy(), inl_o), ms
    with pytest.raises(ValueError):
        do_ransac(g["moving"], g["ransac_f"], 4, 10, 16, 'Rigid')


def test_icp_similar_vs_goldens(O, torch, monkeypatch):
    from platymatch_b200.estimate_transform.perform_icp import perform_icp
    g = load_golden("similar")
    a, resid = perform_icp(g["icp_start"], g["fixed"], 20, 'Similar', verbose=False, return_residuals=True)
    assert np.allclose(a, g["icp_A_fixed"], rtol=1e-6, atol=1e-6)
    a_o, r_o = O.perform_icp(g["icp_start"], g["fixed"], 20, "Similar", return_residuals=True)
    assert np.allclose(resid, r_o, rtol=1e-6, atol=1e-9)
    for env in ("PM_ICP_BRUTE", "PM_ICP_MULTI_LAUNCH"):            # the other two nearest-neighbour code paths
        monkeypatch.setenv(env, "1")
        a2 = perform_icp(g["icp_start"], g["fixed"], 20, 'Similar', verbose=False)
        monkeypatch.delenv(env)
        assert np.allclose(a2, a, rtol=1e-9, atol=1e-9), env


def test_pipeline_similar_vs_oracle(O, torch):
    """Full unsupervised and supervised registration with transform='Similar' against the oracle (corrected Horn)."""
    import platymatch_b200 as pm
    from platymatch_b200.synthetic import make_pair, make_keypoints
    p = make_pair(700, seed=3)
    m, f = p["moving"], p["fixed"]
    idx = [O.ransac_sample_indices(m.shape[1], 4, 400, seed=q) for q in range(4)]
    res = pm.estimate_transform_unsupervised(m, f, ransac_trials=400, sample_indices=idx, transform='Similar')
    ref = O.estimate_transform_unsupervised(m, f, ransac_trials=400, hypotheses=O.HYPOTHESES[:4], sample_indices=idx,
                                            transform="Similar")
    assert res["best"] == ref["best"] and res["inliers"].tolist() == list(ref["inliers"])
    assert np.abs(res["transform"] - ref["transform"]).max() < 1e-4
    lin = res["transform"][:3, :3]
    s2 = lin @ lin.T
    assert np.allclose(s2, s2[0, 0] * np.eye(3), rtol=1e-9, atol=1e-9 * s2[0, 0])          # scale x rotation
    mk, fk = make_keypoints(p, 10, seed=1)
    sup = pm.estimate_transform_supervised(m, f, mk, fk, transform='Similar')
    sup_o = O.estimate_transform_supervised(m, f, mk, fk, transform="Similar")
    assert np.allclose(sup["transform_sc"], sup_o["transform_sc"], rtol=1e-8, atol=1e-8)
    assert np.abs(sup["transform"] - sup_o["transform"]).max() < 1e-4


# ------------------------------------------------------------------------------ rank-deficient affine fits (pinv)
@pytest.mark.parametrize("case", ["one_slice", "tilted_plane", "line", "three_points"])
def test_affine_fit_rank_deficient_equals_pinv(O, torch, case):
    """Keypoints picked in a single z slice (or collinear, or fewer than 4): the reference's fixed_h @ pinv(moving_h)
    is the minimum-norm least-squares affine; the kernels return the same instead of NaN."""
    from platymatch_b200.estimate_transform.find_transform import get_affine_transform
    rng = np.random.default_rng(len(case))
    m = rng.normal(size=(3, 12)) * 60 + np.array([[300.0], [200.0], [150.0]])
    if case == "one_slice":
        m[0] = 37.0
    elif case == "tilted_plane":
        m[2] = 0.3 * m[0] - 0.7 * m[1] + 11.0
    elif case == "line":
        m = np.array([[100.0], [50.0], [20.0]]) + np.outer([1.0, 2.0, -0.5], rng.normal(size=12) * 50)
    else:
        m = m[:, :3]
    f = rng.normal(size=m.shape) * 60 + 250
    ref = O.get_affine_transform(m, f)
    got = get_affine_transform(m, f)
    assert np.all(np.isfinite(got))
    assert np.allclose(got, ref, rtol=1e-7, atol=1e-7 * np.abs(ref).max())


def test_supervised_with_coplanar_keypoints(O, torch):
    """estimate_transform_supervised with all keypoints in one z slice (user clicks in a single plane of the viewer)."""
    import platymatch_b200 as pm
    from platymatch_b200.synthetic import make_pair
    p = make_pair(600, seed=21)
    m, f = p["moving"], p["fixed"]
    z0 = np.median(m[0])
    sel = np.argsort(np.abs(m[0] - z0))[:10]
    mk = m[:, sel].copy()
    mk[0] = z0                                      # exactly coplanar
    fk = f[:, p["gt_fixed_index"][sel]]
    res = pm.estimate_transform_supervised(m, f, mk, fk, icp_iterations=5)
    ref = O.estimate_transform_supervised(m, f, mk, fk, icp_iterations=5)
    assert np.all(np.isfinite(res["transform"]))
    assert np.allclose(res["transform_sc"], ref["transform_sc"], rtol=1e-7, atol=1e-6)
    assert np.abs(res["transform"] - ref["transform"]).max() < 1e-4


def test_flat_clouds_ransac_and_icp(O, torch):
    """2-D data (every nucleus in one plane): every 4-sample is coplanar and the ICP normal equations are singular;
    the reference still answers through pinv.  Inlier counts per trial and the ICP matrix equal the oracle's."""
    from platymatch_b200 import device as D
    from platymatch_b200.estimate_transform.perform_icp import perform_icp
    rng = np.random.default_rng(2)
    k = 400
    m = rng.random((3, k)) * 300
    m[0] = 12.0
    th = 0.3
    A = np.eye(4)
    A[:3, :3] = [[1, 0, 0], [0, np.cos(th), -np.sin(th)], [0, np.sin(th), np.cos(th)]]
    A[:3, 3] = [0.0, 15.0, -7.0]
    f = O.apply_affine_transform(m, A) + rng.normal(0, 0.5, size=(3, k)) * np.array([[0.0], [1.0], [1.0]])
    idx = O.ransac_sample_indices(k, 4, 200, seed=5)
    _, inl_best, inl_o, mats = O.do_ransac(m, f, 4, 200, 3.0, sample_indices=idx, return_all=True)
    md, fd = D.to_device_points(m), D.to_device_points(f)
    a, inl, trial, per = D.ransac(md, fd, 200, 3.0, 4, torch.from_numpy(idx).cuda(), want_per_trial=True)
    assert inl_best > 0.5 * k
    assert np.array_equal(per.cpu().numpy(), inl_o)
    assert np.allclose(a.cpu().numpy().reshape(4, 4)[:3], mats[int(np.argmax(inl_o))][:3], rtol=1e-6, atol=1e-6)
    start = O.apply_affine_transform(m, A) + np.array([[0.0], [2.0], [-1.5]])
    a_icp = perform_icp(start, f, 5, 'Affine', verbose=False)
    a_o = O.perform_icp(start, f, 5)
    assert np.all(np.isfinite(a_icp))
    moved, moved_o = O.apply_affine_transform(start, a_icp), O.apply_affine_transform(start, a_o)
    assert np.abs(moved - moved_o).max() < 1e-6
