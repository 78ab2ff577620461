"""Multi-GPU sharding of the registration path: one process per GPU,
`torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests) for the two real
exchange steps, nothing else.

Three ways the path shards, all with static partitions:

  * specimen pairs      independent registrations -> round-robin over ranks, no data-path collective;
                        one all_gather of the 4x4 results at the end (`register_all_pairs`).
  * hypotheses          the H {cost matrix -> LAP -> RANSAC} chains are independent -> hypothesis q
                        runs on rank q % world; one all_gather of (inliers, 4x4) picks the winner
                        (`reduce_best_hypothesis`), every rank then refines with ICP (replicated,
                        deterministic) so all ranks return the same transform.
  * cost-matrix rows    for large clouds the rows of every cost matrix are computed in row shards
                        (`shard_rows`) and exchanged with one all_gather per matrix (`allgather_rows`)
                        so that the LAP owner holds the full matrix.  The LAP itself does not shard
                        across GPUs ("replicas only"): one matrix, one GPU.

The partition arithmetic and both exchange steps are plain-tensor code (CPU or CUDA), covered by
world_size-2 gloo tests in tests/test_distributed_cpu.py; only `register_pair_sharded` and
`register_all_pairs` touch the CUDA library.
"""
import itertools

import numpy as np


def world_info(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


# ----------------------------------------------------------------------------- partitions
def shard_rows(n_rows, rank, world, align=128):
    """Row range [begin, end) of `rank`: equal shards of ceil(n/world) rounded up to `align`
    (the chi2 kernel's 128-row tile), the tail ranks may be short or empty."""
    per = -(-n_rows // world)
    per = -(-per // align) * align
    begin = min(rank * per, n_rows)
    end = min(begin + per, n_rows)
    return begin, end, per


def hypotheses_for_rank(n_hyp, rank, world):
    """Hypothesis q is owned by rank q % world."""
    return [q for q in range(n_hyp) if q % world == rank]


def all_pairs(n_specimens):
    """Unordered specimen pairs (i < j) in lexicographic order: 66 pairs for 12 specimens."""
    return list(itertools.combinations(range(n_specimens), 2))


def pairs_for_rank(n_specimens, rank, world):
    """Static round-robin of the pair list: pair p goes to rank p % world."""
    return [p for k, p in enumerate(all_pairs(n_specimens)) if k % world == rank]


# ----------------------------------------------------------------------------- exchange steps
def allgather_rows(local_rows, n_rows, per, group=None):
    """All-gather of row shards: local_rows is this rank's [per, ld] block (rows beyond the rank's range
    are padding); returns the full [n_rows, ld] matrix on every rank."""
    import torch
    import torch.distributed as dist
    rank, world = world_info(group)
    if world == 1:
        return local_rows[:n_rows]
    assert local_rows.shape[0] == per
    full = torch.empty((world * per,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(full, local_rows.contiguous(), group=group)
    return full[:n_rows]


def reduce_best_hypothesis(inliers_local, transforms_local, hyp_ids_local, n_hyp, group=None):
    """Best-hypothesis reduction.  Each rank contributes (inliers, 4x4) of the hypotheses it owns; returns
    (inliers[n_hyp], transforms[n_hyp,16], best) on every rank, best = first maximum in hypothesis order
    (np.argmax semantics of reference _dock_widget.py:683-703)."""
    import torch
    import torch.distributed as dist
    rank, world = world_info(group)
    dev = transforms_local.device
    slots = -(-n_hyp // world)
    pack = torch.full((slots, 18), -1.0, dtype=torch.float64, device=dev)
    for k, q in enumerate(hyp_ids_local):
        pack[k, 0] = float(q)
        pack[k, 1] = inliers_local[k].to(torch.float64)
        pack[k, 2:] = transforms_local[k].reshape(16)
    if world > 1:
        gathered = torch.empty((world * slots, 18), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, pack, group=group)
    else:
        gathered = pack
    inliers = torch.zeros(n_hyp, dtype=torch.int64, device=dev)
    transforms = torch.ones((n_hyp, 16), dtype=torch.float64, device=dev)
    g = gathered.cpu()
    for row in g:
        q = int(row[0].item())
        if q >= 0:
            inliers[q] = int(row[1].item())
            transforms[q] = row[2:].to(dev)
    best = int(torch.argmax(inliers).item())
    return inliers, transforms, best


def gather_results(local, n_total, owner_of, group=None):
    """Gather per-item float64 rows (e.g. 16-element transforms) computed by their owner ranks."""
    import torch
    import torch.distributed as dist
    rank, world = world_info(group)
    width = 16
    out = torch.zeros((n_total, width), dtype=torch.float64)
    for k, row in local.items():
        out[k] = torch.as_tensor(row, dtype=torch.float64).reshape(width)
    if world > 1:
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = out.to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)    # rows are disjoint by owner
        out = t.cpu()
    return out


# ----------------------------------------------------------------------------- CUDA paths
def register_pair_sharded(moving, fixed, *, ransac_samples=4, ransac_trials=8000, ransac_error=16, icp_iterations=50,
                          seed=0, hypotheses=None, max_bid_rounds=2048, shard_rows_of_cost=True, group=None):
    """One registration spread over all ranks (configs 2/4 at N > 1): cost-matrix rows of every hypothesis are
    computed in row shards and all-gathered to all ranks; hypothesis q's LAP + RANSAC run on rank q % world;
    (inliers, A) are all-gathered; every rank runs the (deterministic) ICP.  Returns the same dict on every rank."""
    import torch
    from . import device as D, pipeline as P
    rank, world = world_info(group)
    hyps = P.HYPOTHESES_DISTINCT if hypotheses is None else list(hypotheses)
    H = len(hyps)
    need_m, need_f = max(a for a, _ in hyps), max(b for _, b in hyps)
    dm = P.describe_cloud(moving, 1 if need_m == 1 else 2)
    df = P.describe_cloud(fixed, 1 if need_f == 1 else (2 if need_f == 2 else 4))
    n1, n2 = dm.n, df.n
    assert n1 <= n2, "sharded path expects n_moving <= n_fixed (swap the clouds otherwise)"
    ldc = (n2 + 3) // 4 * 4
    mine = hypotheses_for_rank(H, rank, world)
    begin, end, per = shard_rows(n1, rank, world)
    costs = {}
    for q, (a, b) in enumerate(hyps):
        if world > 1 and shard_rows_of_cost:
            local = torch.zeros((per, ldc), dtype=torch.float32, device=dm.pts.device)
            if end > begin:
                D.chi2_cost(dm.operand(a), df.operand(b), out=local, row_begin=begin, row_end=end)
            full = allgather_rows(local, n1, per, group)
            if q in mine:
                costs[q] = full.contiguous()
        elif q in mine:
            costs[q] = D.chi2_cost(dm.operand(a), df.operand(b))
    inl_local, a_local, lap_cost = [], [], {}
    rows = torch.arange(n1, dtype=torch.int32, device=dm.pts.device)
    for q in mine:
        col4row, total, _ = D.lap_solve(costs[q], n1, n2, max_bid_rounds)
        mk = D.gather_points(dm.pts, rows)
        fk = D.gather_points(df.pts, col4row[0].contiguous())
        a, inl, _, _ = D.ransac_affine(mk, fk, int(ransac_trials), float(ransac_error), int(ransac_samples), None,
                                       seed=(int(seed) << 8) + q)
        inl_local.append(inl[0]); a_local.append(a); lap_cost[q] = float(total.item())
    a_stack = torch.stack(a_local) if a_local else torch.empty((0, 16), dtype=torch.float64, device=dm.pts.device)
    inliers, transforms, best = reduce_best_hypothesis(inl_local, a_stack, mine, H, group)
    a_sc = transforms[best].contiguous()
    moved = D.apply_affine(dm.pts, a_sc)
    a_icp, resid, _ = D.icp_affine(moved, df.pts, int(icp_iterations))
    a_final = D.compose(a_icp, a_sc)
    return dict(transform=a_final.cpu().numpy().reshape(4, 4), transform_sc=a_sc.cpu().numpy().reshape(4, 4),
                transform_icp=a_icp.cpu().numpy().reshape(4, 4), inliers=inliers.cpu().numpy(), best=best,
                lap_cost=lap_cost, icp_residuals=resid.cpu().numpy())


def register_all_pairs(specimens, *, group=None, in_flight=3, **kw):
    """Batched all-pairs registration (config 5): descriptors of each specimen are computed once per rank
    that needs them, pairs are sharded round-robin over the ranks, the 4x4 results are gathered on every rank.
    On each rank `in_flight` independent registrations overlap on separate CUDA streams (one host thread each):
    the assignment stage is latency-bound on a few SMs and hides behind the cost-matrix / ICP kernels of the
    other pairs.  specimens: list of 3xN arrays.  Returns (pairs, transforms[n_pairs,4,4])."""
    import threading
    import torch
    from . import pipeline as P
    rank, world = world_info(group)
    pairs = all_pairs(len(specimens))
    mine = [(k, p) for k, p in enumerate(pairs) if k % world == rank]
    desc = {}
    for _, (i, j) in mine:       # variants 1-2 double as the 'moving' sets
        for s in (i, j):
            if s not in desc:
                desc[s] = P.describe_cloud(specimens[s], 4)
    for d in desc.values():      # operands are built lazily: do it here, before the streams fork
        for v in (1, 2, 3, 4):
            d.operand(v)
    torch.cuda.synchronize()
    local, errors = {}, []
    nfl = max(1, min(int(in_flight), len(mine)))
    dev = torch.cuda.current_device()

    def worker(w):
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(torch.cuda.Stream()):
                for k, (i, j) in mine[w::nfl]:      # moving = specimen i, fixed = specimen j (tall problems are transposed inside)
                    res = P.register_described(desc[i], desc[j], seed=k, overlap_hypotheses=(nfl == 1), **kw)
                    local[k] = res["transform"].cpu().numpy().reshape(4, 4)
        except Exception as e:     # surfaced below: a worker must not die silently
            errors.append(e)

    if nfl == 1:
        worker(0)
    else:
        ts = [threading.Thread(target=worker, args=(w,)) for w in range(nfl)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    if errors:
        raise errors[0]
    out = gather_results(local, len(pairs), None, group)
    return pairs, out.numpy().reshape(-1, 4, 4)
