"""Small end-to-end run for compute-sanitizer (memcheck): every kernel family once at small sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import platymatch_b200 as pm
from platymatch_b200.synthetic import make_pair, make_label_volume, make_keypoints
from platymatch_b200.lap import linear_sum_assignment
from platymatch_b200.utils.labels import detections_from_labels
from platymatch_b200.evaluate_metrics import calculate_metrics

p = make_pair(700, seed=3)
res = pm.estimate_transform_unsupervised(p["moving"], p["fixed"], ransac_trials=300, icp_iterations=5, seed=1)
print("unsupervised inliers", res["inliers"].tolist())
res8 = pm.estimate_transform_unsupervised(p["moving"], p["fixed"], ransac_trials=100, icp_iterations=2, as_reference=True)
mk, fk = make_keypoints(p, 10, seed=1)
pm.estimate_transform_supervised(p["moving"], p["fixed"], mk, fk, icp_iterations=3)
rng = np.random.default_rng(0)
for shape, algo, rounds in [((37, 53), 1, 64), ((200, 333), 1, 2048), ((200, 333), 2, 64), ((130, 130), 0, 0), ((5, 3), 1, 8)]:
    r, c = linear_sum_assignment(rng.random(shape), algorithm=algo, max_bid_rounds=rounds)
    print("lap", shape, algo, len(r))
vol = make_label_volume((21, 30, 45), 12, seed=2)
d, s, ids = detections_from_labels(vol, 1.5)
print("labels", d.shape)
n1 = p["moving"].shape[1]
acc, err = calculate_metrics(mk, np.arange(10), p["moving"], np.arange(n1), fk, np.arange(10), p["fixed"],
                             np.arange(700), res["transform_sc"], res["transform_icp"])
print("metrics", acc, err)
torch.cuda.synchronize()
print("done")
