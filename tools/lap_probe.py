"""Developer probe: LAP statistics and timing on the cost matrices of a synthetic pair (GPU)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import platymatch_b200 as pm
from platymatch_b200 import device as D, pipeline as P
from platymatch_b200.synthetic import make_pair

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
algo = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
dropout = float(sys.argv[4]) if len(sys.argv) > 4 else 0.10
seed = int(sys.argv[5]) if len(sys.argv) > 5 else n
p = make_pair(n, seed=seed, dropout=dropout)
dm, df = P.describe_pair(p["moving"], p["fixed"], 1, 4)
n1, n2 = dm.n, df.n
cost = torch.empty((4, n1, (n2 + 3) // 4 * 4), dtype=torch.float32, device="cuda")
for q, (a, b) in enumerate(P.HYPOTHESES_DISTINCT):
    D.chi2_cost(dm.operand(a), df.operand(b), out=cost[q])
names = ["bid_rounds", "rows_after", "augment", "dijkstra", "status", "bids", "refreshes", "retries", "parked",
         "refresh_cyc", "auction_cyc", "bulk_bids", "sap_dense", "-", "-", "-"]
print("n1 x n2 =", n1, "x", n2, flush=True)
for batch in ([0], [1], [2], [3], [0, 1, 2, 3]):
    c = cost[batch].contiguous()
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        col, tot, st = D.lap_solve(c, n1, n2, rounds, algo)
        e1.record()
        torch.cuda.synchronize()
    print("batch", batch, "ms %.3f" % e0.elapsed_time(e1), "total", tot.cpu().numpy())
    for row in st.cpu().numpy():
        print("   ", {k: int(v) for k, v in zip(names, row)})
