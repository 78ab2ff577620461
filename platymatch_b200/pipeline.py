"""Fused entry points restating `EstimateTransform._click_run` (reference
platymatch/_dock_widget.py:526-721) as callables: unsupervised (shape context -> chi^2 cost ->
assignment -> affine RANSAC -> ICP) and keypoint-supervised (LS affine on keypoints -> ICP).

Everything between the host->device copy of the two clouds and the device->host read of the
results stays on the GPU; stage boundaries are stream-ordered kernel launches through the C ABI.
"""
import numpy as np

from . import device as D
from ._lib import LAP_STATS as _LAP_STATS

__all__ = ["HYPOTHESES_REFERENCE", "HYPOTHESES_DISTINCT", "estimate_transform_unsupervised",
           "estimate_transform_supervised", "Descriptors", "describe_cloud", "describe_pair", "register_described"]

# (moving variant a, fixed variant b) of U_ab, in the reference's order 11,12,13,14,21,22,23,24
# (_dock_widget.py:556-602).  sc2/sc3/sc4 are phi-bin permutations of sc, so U21=U12, U22=U11,
# U23=U14, U24=U13 up to bin-edge ties: the last four hypotheses repeat the first four.
HYPOTHESES_REFERENCE = [(1, 1), (1, 2), (1, 3), (1, 4), (2, 1), (2, 2), (2, 3), (2, 4)]
HYPOTHESES_DISTINCT = [(1, 1), (1, 2), (1, 3), (1, 4)]


class Descriptors:
    """Device-resident description of one cloud: points, stats, integer histograms, normalised operands."""

    def __init__(self, pts, stats, mean_dist, counts, dropped, ties):
        self.pts, self.stats, self.mean_dist = pts, stats, mean_dist
        self.counts, self.dropped, self.ties = counts, dropped, ties
        self.n = pts.shape[0]
        self._ops = {}

    def operand(self, variant):
        """chi^2 operand (bin-major float32 histograms + block masks) of `variant` (1-based); either side."""
        if variant not in self._ops:
            self._ops[variant] = D.chi2_operand(self.counts[variant - 1])
        return self._ops[variant]


def describe_cloud(cloud, n_variants, transposed=False):
    """centroid + PCA axis (K0), mean pair distance (K1), shape-context histograms (K2) of one cloud."""
    torch = D._torch()
    if (torch.is_tensor(cloud) and cloud.is_cuda and transposed and cloud.dim() == 2 and cloud.shape[1] == 3
            and cloud.dtype == torch.float64 and cloud.is_contiguous()):
        pts = cloud                       # already the device layout: [N,3] float64
    else:
        pts = D.to_device_points(cloud, transposed=transposed)
    stats = D.cloud_stats(pts)
    md = D.mean_distance(pts)
    counts, dropped, ties = D.shape_context_counts(pts, stats[0:3], stats[3:6], md, n_variants)
    return Descriptors(pts, stats, md, counts, dropped, ties)


_DESC_STREAMS = {}


def describe_pair(moving, fixed, n_variants_moving=1, n_variants_fixed=4, transposed=False, overlap=True):
    """describe_cloud of both clouds; the fixed cloud's kernels run on a side stream next to the moving cloud's
    (they are independent and neither fills the GPU for long), joined before returning."""
    torch = D._torch()
    if not overlap:
        return describe_cloud(moving, n_variants_moving, transposed), describe_cloud(fixed, n_variants_fixed, transposed)
    main = torch.cuda.current_stream()
    key = (torch.cuda.current_device(), main.cuda_stream)
    side = _DESC_STREAMS.get(key)
    if side is None:
        side = _DESC_STREAMS[key] = torch.cuda.Stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        df = describe_cloud(fixed, n_variants_fixed, transposed)
        for v in range(1, n_variants_fixed + 1):
            df.operand(v)
    dm = describe_cloud(moving, n_variants_moving, transposed)
    for v in range(1, n_variants_moving + 1):
        dm.operand(v)
    main.wait_stream(side)
    return dm, df


def _draw_or_take(sample_indices, q):
    if sample_indices is None:
        return None
    torch = D._torch()
    s = sample_indices[q]
    return s if torch.is_tensor(s) else torch.from_numpy(np.ascontiguousarray(s, dtype=np.int32)).cuda()


_HYP_STREAMS = {}


def _hypothesis_streams(device, n):
    """High-priority side streams for the per-hypothesis assignment chains, cached per (device, caller stream):
    concurrent callers (one host thread and stream each) must not share them."""
    torch = D._torch()
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream().cuda_stream)
    pool = _HYP_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device, priority=-1))
    return pool[:n]


def register_described(dm, df, ransac_samples=4, ransac_trials=8000, ransac_error=16, icp_iterations=50,
                       hypotheses=None, seed=0, sample_indices=None, max_bid_rounds=2048, cost_out=None,
                       keep_cost=False, stage_hook=None, overlap_hypotheses=True, transform='Affine'):
    """Cost matrices -> LAP -> RANSAC -> argmax -> ICP for two described clouds (device resident).

    The H hypothesis chains {cost matrix -> assignment -> RANSAC} are independent (reference
    _dock_widget.py:547-675 runs them one after the other).  With `overlap_hypotheses` each chain is enqueued
    on its own CUDA stream: the assignment of one hypothesis is latency-bound on a few SMs and runs underneath
    the compute-bound cost-matrix kernels of the next ones.  Results are identical either way.
    Nothing is read back in between: the reference's np.argmax over the inlier counts (_dock_widget.py:683-703)
    runs on the device (pm_select_best) and ICP takes the winning 4x4 from device memory, so a registration is
    one uninterrupted stream of launches.  Returns a dict of CUDA tensors (`best` is a 1-element int32 tensor).
    """
    torch = D._torch()
    D.transform_code(transform)
    hyps = HYPOTHESES_DISTINCT if hypotheses is None else list(hypotheses)
    n1, n2 = dm.n, df.n
    swap = n1 > n2                      # scipy solves the transpose when there are more rows than columns
    nr, nc = (n2, n1) if swap else (n1, n2)
    ldc = (nc + 3) // 4 * 4
    H = len(hyps)
    dev = dm.pts.device
    cost = cost_out if cost_out is not None else torch.empty((H, nr, ldc), dtype=torch.float32, device=dev)
    ransac_a = torch.empty((H, 16), dtype=torch.float64, device=dev)
    inliers = torch.empty(H, dtype=torch.int32, device=dev)
    col4row = torch.empty((H, nr), dtype=torch.int32, device=dev)
    lap_total = torch.empty(H, dtype=torch.float64, device=dev)
    lap_stats = torch.empty((H, _LAP_STATS), dtype=torch.int64, device=dev)
    rows_arange = torch.arange(nr, dtype=torch.int32, device=dev) if swap else None
    pairs = [None] * H
    ops_m = {a: dm.operand(a) for a, _ in hyps}     # built on the caller's stream, before the chains fork
    ops_f = {b: df.operand(b) for _, b in hyps}

    def cost_matrix(q):
        a, b = hyps[q]
        if swap:
            D.chi2_cost(ops_f[b], ops_m[a], out=cost[q])
        else:
            D.chi2_cost(ops_m[a], ops_f[b], out=cost[q])

    def ransac(q):
        if swap:                        # rows are fixed nuclei; reference order = ascending moving index
            mov_idx, order = torch.sort(col4row[q])
            fix_idx = rows_arange[order].contiguous()
            mk = D.gather_points(dm.pts, mov_idx.contiguous())
        else:                           # moving[:, row_ind] with row_ind = 0..n1-1 is the cloud itself (:622)
            mov_idx, fix_idx = None, col4row[q]
            mk = dm.pts
        fk = D.gather_points(df.pts, fix_idx)
        D.ransac(mk, fk, int(ransac_trials), float(ransac_error), int(ransac_samples), _draw_or_take(sample_indices, q),
                 seed=(int(seed) << 8) + q, transform=transform, out=(ransac_a[q], inliers[q:q + 1]))
        pairs[q] = (mov_idx, fix_idx)

    def lap(q0, q1):
        D.lap_solve(cost[q0:q1], nr, nc, max_bid_rounds,
                    out=(col4row[q0:q1], lap_total[q0:q1], lap_stats[q0:q1]))

    if overlap_hypotheses and stage_hook is None and H > 1:
        # cost matrices stay on the caller's stream (compute-bound, every SM); each assignment + RANSAC chain
        # goes to a high-priority side stream as soon as its matrix is done, so that its few, long-running CTAs
        # are placed ahead of the queued cost-matrix CTAs of the next hypotheses
        main = torch.cuda.current_stream()
        sides = _hypothesis_streams(dev, H)
        for q, side in enumerate(sides):
            cost_matrix(q)
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                lap(q, q + 1)
                ransac(q)
        for side in sides:
            main.wait_stream(side)
    else:
        for q in range(H):
            cost_matrix(q)
        if stage_hook: stage_hook("chi2_cost")
        lap(0, H)
        if stage_hook: stage_hook("lap")
        for q in range(H):
            ransac(q)
        if stage_hook: stage_hook("ransac")
    best, a_sc = D.select_best(inliers, ransac_a)                           # np.argmax, first maximum (:683)
    moved = D.apply_affine(dm.pts, a_sc)                                    # :714
    a_icp, resid, _ = D.icp(moved, df.pts, int(icp_iterations), transform=transform)     # :715-717
    a_final = D.compose(a_icp, a_sc)                                        # icp @ sc (:428)
    if stage_hook: stage_hook("icp")
    out = dict(transform=a_final, transform_sc=a_sc, transform_icp=a_icp, inliers=inliers, ransac_A=ransac_a,
               best=best, hypotheses=hyps, assignments=pairs, lap_cost=lap_total, lap_stats=lap_stats,
               icp_residuals=resid, edge_ties=(dm.ties, df.ties), n_moving=n1)
    if keep_cost:
        out["cost"] = cost
    return out


def _to_host(res):
    """One device -> host pass: every result is copied asynchronously into pinned host memory, then the stream is
    synchronised ONCE (the only host synchronisation of a registration)."""
    torch = D._torch()
    pending = []

    def fetch(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        pending.append(h)
        return h

    staged = {}
    for k, v in res.items():
        if torch.is_tensor(v):
            staged[k] = fetch(v)
        elif k == "assignments":
            staged[k] = [(None if m is None else fetch(m), fetch(f)) for m, f in v]
        elif k == "edge_ties":
            staged[k] = tuple(fetch(t) for t in v)
        else:
            staged[k] = v
    torch.cuda.current_stream().synchronize()
    out = {}
    for k, v in staged.items():
        if k == "assignments":
            n1 = res.get("n_moving")
            v = [((np.arange(len(f), dtype=np.int64) if m is None else m.numpy().astype(np.int64)),
                  f.numpy().astype(np.int64)) for m, f in v]
        elif k == "edge_ties":
            v = tuple(int(t.item()) for t in v)
        elif k == "best":
            v = int(v.item())
        elif torch.is_tensor(v):
            v = v.numpy().copy()
            if k.startswith("transform"):
                v = v.reshape(4, 4)
            elif k == "ransac_A":
                v = v.reshape(-1, 4, 4)
        out[k] = v
    out.pop("n_moving", None)
    return out


def estimate_transform_unsupervised(moving, fixed, *, ransac_samples=4, ransac_trials=8000, ransac_error=16,
                                    icp_iterations=50, transform='Affine', seed=0, sample_indices=None,
                                    as_reference=False, hypotheses=None, max_bid_rounds=2048, keep_cost=False):
    """Unsupervised registration of two nuclei clouds (reference _dock_widget.py:526-721, widget defaults).

    moving, fixed: 3xN (or 4xN) float64 arrays, zyx.  Returns a dict of numpy results:
      transform (= icp @ sc, what `_save_transform` writes), transform_sc, transform_icp,
      inliers[H], ransac_A[H], best, assignments[H] (moving index, fixed index), lap_cost[H],
      icp_residuals, edge_ties.
    `as_reference=True` evaluates all 8 hypotheses like the widget; the default evaluates the 4
    algebraically distinct ones (the other 4 are duplicates up to bin-edge ties).
    transform: 'Affine' or 'Similar' (the widget's choice, :627) for the RANSAC and ICP fits.
    """
    D.transform_code(transform)
    hyps = hypotheses or (HYPOTHESES_REFERENCE if as_reference else HYPOTHESES_DISTINCT)
    need_m = max(a for a, _ in hyps)
    need_f = max(b for _, b in hyps)
    dm, df = describe_pair(moving, fixed, 1 if need_m == 1 else 2, 1 if need_f == 1 else (2 if need_f == 2 else 4))
    res = register_described(dm, df, ransac_samples, ransac_trials, ransac_error, icp_iterations, hyps, seed,
                             sample_indices, max_bid_rounds, keep_cost=keep_cost, transform=transform)
    return _to_host(res)


def estimate_transform_supervised(moving, fixed, moving_keypoints, fixed_keypoints, *, icp_iterations=50,
                                  transform='Affine'):
    """Keypoint-supervised registration (reference _dock_widget.py:707-717): least-squares affine (or Horn
    similarity, :710-711) on the 3xK keypoint pairs, then ICP on the full clouds.  Returns dict(transform,
    transform_sc, transform_icp, icp_residuals) as numpy arrays."""
    D.transform_code(transform)
    m = D.to_device_points(moving)
    f = D.to_device_points(fixed)
    a_sc = D.fit_transform(D.to_device_points(moving_keypoints), D.to_device_points(fixed_keypoints), transform)
    moved = D.apply_affine(m, a_sc)
    a_icp, resid, _ = D.icp(moved, f, int(icp_iterations), transform=transform)
    a_final = D.compose(a_icp, a_sc)
    return _to_host(dict(transform=a_final, transform_sc=a_sc, transform_icp=a_icp, icp_residuals=resid))
