"""Developer probe: every pair of the all-pairs configuration registered alone, one after the other: time, slack
columns of its assignment problems, assignment statistics (`python tools/pairs_probe.py [n] [vary]`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import platymatch_b200 as pm
from platymatch_b200.synthetic import make_specimens

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
vary = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
specs = make_specimens(12, n, seed=0, vary=vary)
pm.estimate_transform_unsupervised(specs[0]["points"], specs[1]["points"], seed=0)
rows = []
for i in range(12):
    for j in range(i + 1, 12):
        a, b = specs[i]["points"], specs[j]["points"]
        if a.shape[1] > b.shape[1]:
            a, b = b, a
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = pm.estimate_transform_unsupervised(a, b, seed=1)
        ms = (time.perf_counter() - t0) * 1e3
        st = np.asarray(res["lap_stats"])
        rows.append((ms, i, j, b.shape[1] - a.shape[1], st[:, 5].tolist(), st[:, 2].tolist(), st[:, 3].tolist(), res["inliers"].tolist()))
for r in sorted(rows, key=lambda r: r[3]):
    print("pair %2d-%2d slack %4d: %8.1f ms  bids %s aug %s dijkstra %s inliers %s" % (r[1], r[2], r[3], r[0], r[4], r[5], r[6], r[7]), flush=True)
print("total %.1f s, slowest %.1f ms" % (sum(r[0] for r in rows) / 1e3, max(r[0] for r in rows)))
