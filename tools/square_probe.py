"""Developer probe: where the time of a registration with few slack columns goes (specimens of the all-pairs row)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import platymatch_b200 as pm
from platymatch_b200 import device as D, pipeline as P
from platymatch_b200.synthetic import make_specimens

specs = make_specimens(12, 8000, seed=0)
pm.estimate_transform_unsupervised(specs[0]["points"], specs[1]["points"], seed=0)
for (i, j) in ((0, 4), (0, 5), (5, 10)):
    a, b = specs[i]["points"], specs[j]["points"]
    if a.shape[1] > b.shape[1]:
        a, b = b, a
    print("pair %d-%d: %d x %d (slack %d)" % (i, j, a.shape[1], b.shape[1], b.shape[1] - a.shape[1]), flush=True)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pm.estimate_transform_unsupervised(a, b, seed=1)
        print("  public API: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
    m, f = D.to_device_points(a), D.to_device_points(b)
    for overlap in (True, False):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dm, df = P.describe_pair(m, f, 1, 4, transposed=True)
        res = P.register_described(dm, df, seed=1, overlap_hypotheses=overlap)
        torch.cuda.synchronize()
        print("  register_described overlap=%s: %.1f ms" % (overlap, (time.perf_counter() - t0) * 1e3), flush=True)
    marks = []
    def hook(name):
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
    hook("start")
    dm, df = P.describe_pair(m, f, 1, 4, transposed=True); hook("describe")
    res = P.register_described(dm, df, seed=1, stage_hook=hook, keep_cost=True)
    torch.cuda.synchronize()
    print("  stages:", {n1: round(a_.elapsed_time(b_), 2) for (n0, a_), (n1, b_) in zip(marks[:-1], marks[1:])}, flush=True)
    cost = res["cost"]
    n1, n2 = dm.n, df.n
    for q in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.lap_solve(cost[q:q + 1], n1, n2); e1.record(); torch.cuda.synchronize()
        print("  lap alone hyp %d: %.1f ms" % (q, e0.elapsed_time(e1)), flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); D.lap_solve(cost, n1, n2); e1.record(); torch.cuda.synchronize()
    print("  lap batch of 4: %.1f ms" % e0.elapsed_time(e1), flush=True)
