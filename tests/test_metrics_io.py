"""SURVEY §8(f) rows 3 and 4 — EvaluateMetrics numbers (reference _dock_widget.py:1030-1080) and the file formats
(utils.py:19-43, _dock_widget.py:426-432)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_detection_and_transform_files_round_trip(tmp_path):
    from platymatch_b200.utils.io import load_detections, load_transform, save_transform
    g = np.load(os.path.join(ROOT, "tests", "golden", "asset02.npz"))
    zyx = g["moving"][:3]                                   # the reference's asset 02, already flipped to z y x
    ids = np.arange(100, 100 + zyx.shape[1])
    rows = np.column_stack([ids, zyx[2], zyx[1], zyx[0]])   # file order: id x y z  (utils.py:27-33)
    p = tmp_path / "det.csv"
    np.savetxt(p, rows, delimiter=" ", fmt="%.5f")
    det, got_ids = load_detections(p)
    assert det.shape == zyx.shape and np.allclose(det, zyx, atol=1e-5) and np.array_equal(got_ids, ids)
    with open(tmp_path / "det_h.csv", "w") as f:            # header row skipped (utils.py:23)
        f.write("id x y z\n" + open(p).read())
    det_h, _ = load_detections(tmp_path / "det_h.csv", header=True)
    assert np.array_equal(det_h, det)
    rows_izyx = np.column_stack([ids, zyx[0], zyx[1], zyx[2], np.full(len(ids), 3.0)])      # id z y x r: no flip
    np.savetxt(tmp_path / "izyx.csv", rows_izyx, delimiter=" ", fmt="%.5f")
    det_i, _ = load_detections(tmp_path / "izyx.csv", izyx=True)
    assert np.allclose(det_i, zyx, atol=1e-5)
    a_sc = np.array([[0.9, 0.1, 0, 12.3456], [-0.1, 0.9, 0, -7.0004], [0, 0, 1.1, 3.0], [0, 0, 0, 1.0]])
    a_icp = np.array([[1, 0, 0, 0.5], [0, 1, 0, 0.25], [0, 0, 1, -1.0], [0, 0, 0, 1.0]])
    saved = save_transform(tmp_path / "t.txt", a_icp, a_sc)
    assert np.array_equal(saved, a_icp @ a_sc)
    assert open(tmp_path / "t.txt").read().splitlines()[0] == "0.900 0.100 0.000 12.846"   # fmt='%1.3f' (:432)
    back = load_transform(tmp_path / "t.txt")
    assert back.shape == (4, 4) and np.allclose(back, a_icp @ a_sc, atol=5e-4)
    np.savetxt(tmp_path / "bad.txt", np.eye(3), delimiter=" ")
    with pytest.raises(AssertionError):
        load_transform(tmp_path / "bad.txt")


def _metric_case(tag):
    g = np.load(os.path.join(ROOT, "tests", "golden", "rows.npz"))
    k = lambda name: g["met_%s_%s" % (tag, name)]
    args = (k("mk"), k("kp_ids"), k("moving"), k("mids"), k("fk"), k("kp_ids"), k("fixed"), k("fids"), k("t1"), k("t2"))
    return args, float(k("accuracy")), float(k("error"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_metrics_match_reference_widget(tag):
    """tests/golden/rows.npz: the two numbers the reference's own `EvaluateMetrics._calculate_metrics` wrote into its
    line edits ('{:.3f}'), run headless on a fake widget by oracle/make_golden_rows.py."""
    import oracle as O
    args, acc, err = _metric_case(tag)
    racc, rerr = O.calculate_metrics(*args)
    assert float("%.3f" % racc) == acc and float("%.3f" % rerr) == err


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_metrics_match_reference_widget(tag):
    from platymatch_b200.evaluate_metrics import calculate_metrics
    args, acc, err = _metric_case(tag)
    gacc, gerr = calculate_metrics(*args)
    assert float("%.3f" % gacc) == acc and float("%.3f" % gerr) == err


@pytest.mark.gpu
def test_cdist_matches_scipy():
    import torch
    from scipy.spatial.distance import cdist
    from platymatch_b200 import device as D
    rng = np.random.default_rng(0)
    for n1, n2 in [(1, 1), (7, 130), (65, 64), (300, 517)]:
        a, b = rng.normal(size=(n1, 3)) * 100, rng.normal(size=(n2, 3)) * 100
        got = D.cdist(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()[:, :n2]
        assert np.array_equal(got, cdist(a, b).astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("n,transform", [(400, "gt"), (1500, "estimated"), (400, "identity")])
def test_calculate_metrics_vs_oracle(n, transform):
    import oracle as O
    import platymatch_b200 as pm
    from platymatch_b200.evaluate_metrics import calculate_metrics
    from platymatch_b200.synthetic import make_pair, make_keypoints
    p = make_pair(n, seed=n + 1)
    rng = np.random.default_rng(n)
    sel = rng.choice(p["moving"].shape[1], 12, replace=False)
    mk = p["moving"][:, sel] + rng.normal(0, 0.5, size=(3, 12))
    fk = p["fixed"][:, p["gt_fixed_index"][sel]] + rng.normal(0, 0.5, size=(3, 12))
    kp_ids = np.arange(1, 13)
    fixed_kp_ids = kp_ids.copy()
    mids, fids = 1000 + np.arange(p["moving"].shape[1]), 5000 + np.arange(p["fixed"].shape[1])
    if transform == "gt":
        t1, t2 = p["A_gt"], None
    elif transform == "identity":
        t1, t2 = np.eye(4), np.eye(4)
    else:
        res = pm.estimate_transform_unsupervised(p["moving"], p["fixed"], ransac_trials=1000, seed=1)
        t1, t2 = res["transform_sc"], res["transform_icp"]
    acc, err = calculate_metrics(mk, kp_ids, p["moving"], mids, fk, fixed_kp_ids, p["fixed"], fids, t1, t2)
    racc, rerr = O.calculate_metrics(mk, kp_ids, p["moving"], mids, fk, fixed_kp_ids, p["fixed"], fids, t1, t2)
    assert acc == racc, (acc, racc)
    assert err == pytest.approx(rerr, rel=1e-9)
    if transform != "identity":
        assert acc > 0.7 and err < 5.0
