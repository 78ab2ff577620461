"""Multi-GPU check (torchrun, NCCL; also runs on one GPU): the sharded registration of one pair — cost rows through
the owner's peer-mapped window and through dist.gather — and the all-pairs batch (specimens of UNEQUAL sizes, three
registrations in flight per GPU, shared work counter and static schedule) give the same results as the single-GPU
pipeline; times the 20k x 20k configuration when asked.  Prints DIST_CHECK_OK on rank 0 when everything holds.

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
from platymatch_b200.synthetic import make_pair, make_specimens

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def say(*a):
    if rank == 0:
        print(*a, flush=True)


p = make_pair(3000, seed=5)
single = pm.estimate_transform_unsupervised(p["moving"], p["fixed"], ransac_trials=2000, seed=3)
for peer in ((True, False) if world > 1 else (False,)):
    shard = PD.register_pair_sharded(p["moving"], p["fixed"], ransac_trials=2000, seed=3, peer_stores=peer)
    err = float(np.abs(single["transform"] - shard["transform"]).max())
    assert np.array_equal(single["inliers"], shard["inliers"]), (single["inliers"], shard["inliers"])
    for q, c in shard["lap_cost"].items():
        assert abs(c - single["lap_cost"][q]) <= 1e-12 * abs(c), (q, c, single["lap_cost"][q])
    assert err < 1e-9, err
    assert shard["peer_stores"] == (peer and world > 1)
    say("sharded pair (peer stores %s) == single GPU: max|dT| %.2e, inliers %s, %d bytes sent to other owners by rank 0"
        % (peer, err, shard["inliers"].tolist(), shard["exchanged_bytes"]))

# all pairs: 5 specimens whose sizes differ by up to 12 % (tall and wide problems, different shared-memory sizes of
# the assignment kernels in flight at the same time), three registrations in flight per GPU
specs = [s["points"] for s in make_specimens(5, 1400, seed=3, vary=0.12)]
ref = None
for dynamic in (True, False):
    st = {}
    pairs, T = PD.register_all_pairs(specs, ransac_trials=500, in_flight=3, dynamic=dynamic, stats=st)
    if ref is None:
        ref = [pm.estimate_transform_unsupervised(specs[i], specs[j], ransac_trials=500, seed=k)["transform"]
               for k, (i, j) in enumerate(pairs)]
    worst = max(float(np.abs(a - b).max()) for a, b in zip(ref, T))
    assert worst < 1e-9, worst
    say("all-pairs batch (%d pairs, sizes %s, dynamic %s) == single GPU: max|dT| %.2e; rank 0 did %d pairs"
        % (len(pairs), [s.shape[1] for s in specs], dynamic, worst, st["pairs_done"]))

if len(sys.argv) > 1:
    n = int(sys.argv[1])
    big = make_pair(n, seed=2)          # (seed = n happens to be a pair the method cannot register)
    for peer in ((True, False) if world > 1 else (False,)):
        for rep in range(3):
            tm = {}
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            res = PD.register_pair_sharded(big["moving"], big["fixed"], seed=1, peer_stores=peer, timings=tm)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            dt = time.perf_counter() - t0
        e = float(np.abs(res["transform"] - big["A_gt"]).max())
        say("n=%d world=%d peer stores %s: %.1f ms per registration (host clock, incl. H2D), stages %s, inliers %s, |T - A_gt| %.3g"
            % (n, world, peer, dt * 1e3, {k: round(v, 1) for k, v in tm.items()}, res["inliers"].tolist(), e))
PD.close_windows()
say("DIST_CHECK_OK")
if world > 1:
    dist.destroy_process_group()
