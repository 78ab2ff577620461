"""Developer probe: assignment time against the number of slack columns (n_fixed - n_moving), all 4 hypothesis
matrices of a synthetic pair solved as one batch.  `python tools/lap_slack.py n slack [slack ...]`"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from platymatch_b200 import device as D, pipeline as P
from platymatch_b200.synthetic import make_pair

n = int(sys.argv[1])
for slack in [int(a) for a in sys.argv[2:]]:
    p = make_pair(n, seed=n, dropout=slack / n)
    dm, df = P.describe_pair(p["moving"], p["fixed"], 1, 4)
    n1, n2 = dm.n, df.n
    cost = torch.empty((4, n1, (n2 + 3) // 4 * 4), dtype=torch.float32, device="cuda")
    for q, (a, b) in enumerate(P.HYPOTHESES_DISTINCT):
        D.chi2_cost(dm.operand(a), df.operand(b), out=cost[q])
    for q in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        col, tot, st = D.lap_solve(cost[q:q + 1], n1, n2)
        e1.record()
        torch.cuda.synchronize()
        s = st[0].cpu().numpy()
        print("n %d slack %d hyp %d: %.2f ms  bids %d (bulk %d) refreshes %d parked %d | augment %d dijkstra %d dense %d | total %.6f"
              % (n, n2 - n1, q, e0.elapsed_time(e1), s[5], s[11], s[6], s[8], s[2], s[3], s[12], tot.item()), flush=True)
