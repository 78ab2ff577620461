"""Developer probe: run-to-run variation of the LAP time on the bench's four pairs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from platymatch_b200 import device as D, pipeline as P
from platymatch_b200.synthetic import make_pair
n = 8000
for slot in range(4):
    p = make_pair(n, seed=n + 97 * slot)
    dm, df = P.describe_cloud(p["moving"], 1), P.describe_cloud(p["fixed"], 4)
    n1, n2 = dm.n, df.n
    cost = torch.empty((4, n1, n2), dtype=torch.float32, device="cuda")
    for q, (a, b) in enumerate(P.HYPOTHESES_DISTINCT):
        D.chi2_cost(dm.operand(a), df.operand(b), out=cost[q])
    ts = []
    for rep in range(12):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        col, tot, st = D.lap_solve(cost, n1, n2)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        if ts[-1] > 1.5 * min(ts) and rep > 0:
            print("   slow rep", rep, "%.2f ms" % ts[-1], "bids", st[:, 5].tolist(), "bulk", st[:, 11].tolist(), "augment", st[:, 2].tolist(),
                  "dijkstra", st[:, 3].tolist(), "auction_cyc", st[:, 10].tolist(), "dense", st[:, 12].tolist())
    print("slot", slot, "ms:", " ".join("%.1f" % t for t in ts))
    print("   last rep: bids", st[:, 5].tolist(), "bulk", st[:, 11].tolist(), "refresh", st[:, 6].tolist(), "parked", st[:, 8].tolist(),
          "augment", st[:, 2].tolist(), "dijkstra", st[:, 3].tolist(), "tail_cyc", st[:, 10].tolist(), "dense", st[:, 12].tolist())
    for q in range(4):      # each matrix alone
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.lap_solve(cost[q:q + 1], n1, n2); e1.record(); torch.cuda.synchronize()
        print("   matrix %d alone: %.2f ms" % (q, e0.elapsed_time(e1)))
