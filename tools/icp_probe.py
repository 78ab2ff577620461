"""Developer probe: ICP time against the cloud size, for the three nearest-neighbour code paths
(`python tools/icp_probe.py 8000 20000`; PM_ICP_MULTI_LAUNCH=1 / PM_ICP_BRUTE=1 select the fallbacks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from platymatch_b200 import device as D
from platymatch_b200.synthetic import make_pair

for n in [int(a) for a in sys.argv[1:]]:
    p = make_pair(n, seed=2)
    a = p["A_gt"].copy()
    a[:3, 3] += 3.0                                   # a good but not perfect initial transform, like RANSAC's
    m, f = D.to_device_points(p["moving"]), D.to_device_points(p["fixed"])
    moved = D.apply_affine(m, torch.from_numpy(a.reshape(16)).cuda())
    for mode in ("", "PM_ICP_MULTI_LAUNCH", "PM_ICP_BRUTE"):
        if mode:
            os.environ[mode] = "1"
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a_icp, resid, _ = D.icp(moved, f, 50)
            e1.record()
            torch.cuda.synchronize()
        if mode:
            del os.environ[mode]
        r = resid.cpu().numpy()
        print("n %d  %-20s %.3f ms for 50 iterations; residual %.3f -> %.3f" % (n, mode or "persistent", e0.elapsed_time(e1), r[0], r[-1]), flush=True)
