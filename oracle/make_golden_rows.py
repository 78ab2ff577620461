"""Golden vectors for the widened rows, dumped from the UNMODIFIED reference.

The reference offers no callables for these rows: the label-image loops are inline in
`EstimateTransform._click_run` (platymatch/_dock_widget.py:497-521) and the metrics live in the Qt method
`EvaluateMetrics._calculate_metrics` (:1030-1080).  Both only touch attributes of `self`, so they are executed
here headless on a fake `self` (oracle/ref_shim.py stubs Qt / napari): `_click_run` is advanced through its label
loops until it has stored `self.moving_detections` / `self.fixed_detections`; `_calculate_metrics` is called
unbound and the two numbers are read back from the fake line edits.

    python oracle/make_golden_rows.py        # container only (needs /root/reference); writes tests/golden/rows.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

ref_shim.install()
import platymatch._dock_widget as W  # noqa: E402
from platymatch_b200.synthetic import make_label_volume, make_pair  # noqa: E402


class LineEdit:
    def __init__(self, text=""):
        self._t = text

    def text(self):
        return self._t

    def setText(self, t):
        self._t = t


class Check:
    def __init__(self, v):
        self.v = v

    def isChecked(self):
        return self.v


class Combo:
    def __init__(self, i):
        self.i = i

    def currentIndex(self):
        return self.i


def reference_label_detections(moving_vol, fixed_vol, anisotropy_m, anisotropy_f):
    """Drive the reference's own _click_run through its label loops (:497-521)."""
    layers = [types.SimpleNamespace(data=moving_vol), types.SimpleNamespace(data=fixed_vol)]
    me = types.SimpleNamespace(
        run_button=types.SimpleNamespace(setStyleSheet=lambda *a: None),
        csv_checkbox=Check(False), viewer=types.SimpleNamespace(layers=layers),
        moving_image_combobox=Combo(0), fixed_image_combobox=Combo(1),
        moving_image_anisotropy_line=LineEdit(str(anisotropy_m)), fixed_image_anisotropy_line=LineEdit(str(anisotropy_f)),
        shape_context_checkbox=Check(False), pca_checkbox=Check(False))
    gen = W.EstimateTransform._click_run(me)
    try:
        for _ in gen:
            pass
    except Exception as e:      # whatever follows the detections in the widget is not part of this row
        print("   (_click_run stopped after the label loops: %s: %s)" % (type(e).__name__, e))
    return np.asarray(me.moving_detections), np.asarray(me.fixed_detections)


def reference_metrics(mk, mk_ids, md, mids, fk, fk_ids, fd, fids, t1, t2):
    me = types.SimpleNamespace(
        moving_keypoints=mk, moving_keypoint_ids=mk_ids, moving_detections=md, moving_ids=mids,
        fixed_keypoints=fk, fixed_keypoint_ids=fk_ids, fixed_detections=fd, fixed_ids=fids,
        transform_matrix_1=t1, transform_matrix_2=t2, transform_matrix_combined=np.matmul(t2, t1),
        matching_accuracy_linedit=LineEdit(), avg_registration_error_lineedit=LineEdit())
    W.EvaluateMetrics._calculate_metrics(me)
    return float(me.matching_accuracy_linedit.text()), float(me.avg_registration_error_lineedit.text())


def main():
    out = {}
    # ---- row 2
    mv = make_label_volume((20, 28, 36), 14, radius=(2.0, 4.0), seed=3, dtype=np.int32)
    fv = make_label_volume((24, 26, 30), 11, radius=(2.0, 4.5), seed=4, dtype=np.uint16, sparse_ids=True)
    md, fd = reference_label_detections(mv, fv, 1.0, 2.5)
    out.update(label_moving_vol=mv, label_fixed_vol=fv, label_moving_det=md, label_fixed_det=fd)
    print("row 2: detections", md.shape, fd.shape)
    # ---- row 3
    for tag, n, noise in (("a", 150, 0.5), ("b", 260, 3.0)):
        p = make_pair(n, seed=40 + n)
        rng = np.random.default_rng(n)
        sel = rng.choice(p["moving"].shape[1], 9, replace=False)
        mk = p["moving"][:, sel] + rng.normal(0, noise, size=(3, 9))
        fk = p["fixed"][:, p["gt_fixed_index"][sel]] + rng.normal(0, noise, size=(3, 9))
        kp_ids = np.arange(1, 10)
        mids, fids = 100 + np.arange(p["moving"].shape[1]), 700 + np.arange(p["fixed"].shape[1])
        t1 = p["A_gt"].copy()
        t2 = np.eye(4); t2[:3, 3] = [0.3, -0.2, 0.1]
        acc, err = reference_metrics(mk, kp_ids, p["moving"], mids, fk, kp_ids.copy(), p["fixed"], fids, t1, t2)
        print("row 3 case %s: accuracy %.3f error %.3f" % (tag, acc, err))
        out.update({"met_%s_mk" % tag: mk, "met_%s_fk" % tag: fk, "met_%s_kp_ids" % tag: kp_ids,
                    "met_%s_moving" % tag: p["moving"], "met_%s_fixed" % tag: p["fixed"], "met_%s_mids" % tag: mids,
                    "met_%s_fids" % tag: fids, "met_%s_t1" % tag: t1, "met_%s_t2" % tag: t2,
                    "met_%s_accuracy" % tag: acc, "met_%s_error" % tag: err})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rows.npz"), **out)
    print("wrote tests/golden/rows.npz")


if __name__ == "__main__":
    main()
