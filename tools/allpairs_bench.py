"""BASELINE config 5: batched all-pairs registration of 12 synthetic specimens (66 pairs, 8k nuclei each), pairs
sharded over the ranks (torchrun, NCCL; also runs on one GPU).  Prints pairs/s and how many pairs were recovered
(ground truth: specimens are affine views of one atlas, so the transform of pair (i, j) is A_j A_i^-1).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/allpairs_bench.py [in_flight]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from platymatch_b200 import distributed as PD
from platymatch_b200.synthetic import make_specimens

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
in_flight = int(sys.argv[1]) if len(sys.argv) > 1 else 3
specs = make_specimens(12, 8000, seed=0)
clouds = [s["points"] for s in specs]
PD.register_all_pairs(clouds[:3], in_flight=in_flight)            # warm-up (3 pairs)
best = None
for rep in range(2):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pairs, T = PD.register_all_pairs(clouds, in_flight=in_flight)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
if rank == 0:
    ok, errs = 0, []
    for (i, j), t in zip(pairs, T):
        gt = specs[j]["A"] @ np.linalg.inv(specs[i]["A"])
        pts = np.vstack([specs[i]["points"][:, :500], np.ones((1, 500))])
        e = np.median(np.linalg.norm((t @ pts)[:3] - (gt @ pts)[:3], axis=0))
        errs.append(e)
        ok += e < 4.0
    print("config 5: %d pairs, world %d, %d in flight per GPU: %.3f s (host clock, descriptors of the 12 specimens and H2D "
          "included) = %.1f pairs/s; %d/%d pairs recovered (median point error < 4 px; worst recovered %.2f px)"
          % (len(pairs), world, in_flight, best, len(pairs) / best, ok, len(pairs), max([e for e in errs if e < 4.0], default=float('nan'))), flush=True)
if world > 1:
    dist.destroy_process_group()
