"""Multi-GPU check (run under torchrun, NCCL): the sharded registration of one pair and the all-pairs batch
give the same transforms as the single-GPU pipeline; times the 20k x 20k configuration when asked.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py [n_big]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import platymatch_b200 as pm
from platymatch_b200 import distributed as PD
from platymatch_b200.synthetic import make_pair

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

p = make_pair(3000, seed=5)
single = pm.estimate_transform_unsupervised(p["moving"], p["fixed"], ransac_trials=2000, seed=3)
shard = PD.register_pair_sharded(p["moving"], p["fixed"], ransac_trials=2000, seed=3)
err = float(np.abs(single["transform"] - shard["transform"]).max())
assert np.array_equal(single["inliers"], shard["inliers"]), (single["inliers"], shard["inliers"])
assert np.allclose([shard["lap_cost"][q] for q in sorted(shard["lap_cost"])],
                   [single["lap_cost"][q] for q in sorted(shard["lap_cost"])], rtol=1e-12)
assert err < 1e-9, err
gt = float(np.abs(shard["transform"] - p["A_gt"]).max()) if "A_gt" in p else float("nan")
if rank == 0:
    print("sharded pair == single GPU: max|dT| %.2e, inliers %s, |T - A_gt| %.3g" % (err, shard["inliers"].tolist(), gt), flush=True)

specs = [make_pair(900, seed=20 + s)["fixed"] for s in range(4)]
pairs, T = PD.register_all_pairs(specs, ransac_trials=500)
if rank == 0:
    ref = [pm.estimate_transform_unsupervised(specs[i], specs[j], ransac_trials=500, seed=k)["transform"]
           for k, (i, j) in enumerate(pairs)]
    print("all-pairs batch (%d pairs) == single GPU: max|dT| %.2e" % (len(pairs), max(np.abs(a - b).max() for a, b in zip(ref, T))), flush=True)

if len(sys.argv) > 1:
    n = int(sys.argv[1])
    big = make_pair(n, seed=2)          # (seed = n happens to be a pair the method cannot register)
    for rep in range(2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = PD.register_pair_sharded(big["moving"], big["fixed"], seed=1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
    if rank == 0:
        e = float(np.abs(res["transform"] - big["A_gt"]).max()) if "A_gt" in big else float("nan")
        print("n=%d world=%d: %.1f ms per registration (host clock, incl. H2D), inliers %s, |T - A_gt| %.3g"
              % (n, world, dt * 1e3, res["inliers"].tolist(), e), flush=True)
if world > 1:
    dist.destroy_process_group()
