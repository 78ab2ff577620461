"""Developer probe: registration quality and LAP optimality (vs scipy on the same float32 matrix) at large N."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import platymatch_b200 as pm
from platymatch_b200.synthetic import make_pair

for n in [int(a) for a in sys.argv[1:]]:
    p = make_pair(n, seed=n)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = pm.estimate_transform_unsupervised(p["moving"], p["fixed"], seed=1, keep_cost=True)
    dt = time.perf_counter() - t0
    best = res["best"]
    n1 = p["moving"].shape[1]
    match = [float(np.mean(res["assignments"][q][1] == p["gt_fixed_index"])) for q in range(4)]
    moved = (res["transform"] @ np.vstack([p["moving"], np.ones((1, n1))]))[:3]
    err = np.linalg.norm(moved - p["fixed"][:, p["gt_fixed_index"]], axis=0)
    print("n=%d  %.0f ms  inliers %s best %d  correct-match fraction per hypothesis %s  median err %.2f px"
          % (n, dt * 1e3, res["inliers"].tolist(), best, ["%.3f" % m for m in match], np.median(err)), flush=True)
    st = res["lap_stats"]
    print("   lap stats (bids, refreshes, parked, augment, dijkstra):", [[int(s[5]), int(s[6]), int(s[8]), int(s[2]), int(s[3])] for s in st], flush=True)
    if os.environ.get("SCIPY_CHECK"):
        from scipy.optimize import linear_sum_assignment
        for q in range(int(os.environ["SCIPY_CHECK"])):
            c = res["cost"][q][:, :n].astype(np.float64)
            t0 = time.perf_counter(); r, cc = linear_sum_assignment(c); ts = time.perf_counter() - t0
            print("   hypothesis %d: scipy %.1f s  cost %.10f   gpu cost %.10f   same assignment %s" %
                  (q, ts, c[r, cc].sum(), res["lap_cost"][q], bool(np.array_equal(cc, res["assignments"][q][1]))), flush=True)
