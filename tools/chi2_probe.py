"""Developer probe: the cost-matrix kernel alone on the headline shell pair and on a filled-ellipsoid pair (dense
histograms), for ncu (`ncu --set full -k regex:pm_chi2_kernel ... python tools/chi2_probe.py`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from platymatch_b200 import device as D, pipeline as P
from platymatch_b200.synthetic import make_pair

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
for filled in (False, True):
    p = make_pair(n, filled=filled)
    dm, df = P.describe_pair(p["moving"], p["fixed"], 1, 4)
    out = torch.empty((dm.n, (df.n + 3) // 4 * 4), dtype=torch.float32, device="cuda")
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D.chi2_cost(dm.operand(1), df.operand(1), out=out)
        e1.record()
        torch.cuda.synchronize()
    print("filled" if filled else "shell", "%d x %d: %.4f ms" % (dm.n, df.n, e0.elapsed_time(e1)), flush=True)
