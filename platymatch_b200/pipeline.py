"""Fused entry points restating `EstimateTransform._click_run` (reference
platymatch/_dock_widget.py:526-721) as callables: unsupervised (shape context -> chi^2 cost ->
assignment -> affine RANSAC -> ICP) and keypoint-supervised (LS affine on keypoints -> ICP).

Everything between the host->device copy of the two clouds and the device->host read of the
results stays on the GPU; stage boundaries are stream-ordered kernel launches through the C ABI.
"""
import numpy as np

from . import device as D

__all__ = ["HYPOTHESES_REFERENCE", "HYPOTHESES_DISTINCT", "estimate_transform_unsupervised",
           "estimate_transform_supervised", "Descriptors", "describe_cloud", "register_described"]

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


def _draw_or_take(sample_indices, q):
    if sample_indices is None:
        return None
    torch = D._torch()
    s = sample_indices[q]
    return s if torch.is_tensor(s) else torch.from_numpy(np.ascontiguousarray(s, dtype=np.int32)).cuda()


def register_described(dm, df, ransac_samples=4, ransac_trials=8000, ransac_error=16, icp_iterations=50,
                       hypotheses=None, seed=0, sample_indices=None, max_bid_rounds=2048, cost_out=None,
                       keep_cost=False, stage_hook=None):
    """Cost matrices -> LAP -> RANSAC -> argmax -> ICP for two described clouds (device resident).

    Returns a dict of CUDA tensors / python scalars; only `best` (one int) is read back in between
    (the reference's np.argmax over the inlier counts, _dock_widget.py:683-703).
    """
    torch = D._torch()
    hyps = HYPOTHESES_DISTINCT if hypotheses is None else list(hypotheses)
    n1, n2 = dm.n, df.n
    swap = n1 > n2                      # scipy solves the transpose when there are more rows than columns
    nr, nc = (n2, n1) if swap else (n1, n2)
    ldc = (nc + 3) // 4 * 4
    H = len(hyps)
    cost = cost_out if cost_out is not None else torch.empty((H, nr, ldc), dtype=torch.float32, device=dm.pts.device)
    for q, (a, b) in enumerate(hyps):
        if swap:
            D.chi2_cost(df.operand(b), dm.operand(a), out=cost[q])
        else:
            D.chi2_cost(dm.operand(a), df.operand(b), out=cost[q])
    if stage_hook: stage_hook("chi2_cost")
    col4row, lap_total, lap_stats = D.lap_solve(cost, nr, nc, max_bid_rounds)
    if stage_hook: stage_hook("lap")
    ransac_a = torch.empty((H, 16), dtype=torch.float64, device=dm.pts.device)
    inliers = torch.empty(H, dtype=torch.int32, device=dm.pts.device)
    pairs = []
    rows_arange = torch.arange(nr, dtype=torch.int32, device=dm.pts.device)
    for q in range(H):
        if swap:                        # rows are fixed nuclei; reference order = ascending moving index
            mov_idx, order = torch.sort(col4row[q])
            fix_idx = rows_arange[order]
            mov_idx = mov_idx.contiguous()
        else:
            mov_idx, fix_idx = rows_arange, col4row[q]
        mk = D.gather_points(dm.pts, mov_idx)
        fk = D.gather_points(df.pts, fix_idx.contiguous())
        a, inl, _, _ = D.ransac_affine(mk, fk, int(ransac_trials), float(ransac_error), int(ransac_samples),
                                       _draw_or_take(sample_indices, q), seed=(int(seed) << 8) + q)
        ransac_a[q] = a
        inliers[q:q + 1] = inl
        pairs.append((mov_idx, fix_idx))
    if stage_hook: stage_hook("ransac")
    best = int(torch.argmax(inliers).item())     # first maximum, as np.argmax (_dock_widget.py:683)
    a_sc = ransac_a[best].contiguous()
    moved = D.apply_affine(dm.pts, a_sc)                                   # :714
    a_icp, resid, _ = D.icp_affine(moved, df.pts, int(icp_iterations))     # :715-717
    a_final = D.compose(a_icp, a_sc)                                       # icp @ sc (:428)
    if stage_hook: stage_hook("icp")
    out = dict(transform=a_final, transform_sc=a_sc, transform_icp=a_icp, inliers=inliers, ransac_A=ransac_a,
               best=best, hypotheses=hyps, assignments=pairs, lap_cost=lap_total, lap_stats=lap_stats,
               icp_residuals=resid, edge_ties=(dm.ties, df.ties))
    if keep_cost:
        out["cost"] = cost
    return out


def _to_host(res):
    torch = D._torch()
    out = {}
    for k, v in res.items():
        if torch.is_tensor(v):
            v = v.cpu().numpy()
            if k.startswith("transform"):
                v = v.reshape(4, 4)
            elif k == "ransac_A":
                v = v.reshape(-1, 4, 4)
        elif k == "assignments":
            v = [(m.cpu().numpy().astype(np.int64), f.cpu().numpy().astype(np.int64)) for m, f in v]
        elif k == "edge_ties":
            v = tuple(int(t.item()) for t in v)
        out[k] = v
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
    """
    if transform != 'Affine':
        raise NotImplementedError("transform='Similar' is a SURVEY §8(f) 'next' row; only 'Affine' is built")
    hyps = hypotheses or (HYPOTHESES_REFERENCE if as_reference else HYPOTHESES_DISTINCT)
    need_m = max(a for a, _ in hyps)
    need_f = max(b for _, b in hyps)
    dm = describe_cloud(moving, 1 if need_m == 1 else 2)
    df = describe_cloud(fixed, 1 if need_f == 1 else (2 if need_f == 2 else 4))
    res = register_described(dm, df, ransac_samples, ransac_trials, ransac_error, icp_iterations, hyps, seed,
                             sample_indices, max_bid_rounds, keep_cost=keep_cost)
    return _to_host(res)


def estimate_transform_supervised(moving, fixed, moving_keypoints, fixed_keypoints, *, icp_iterations=50,
                                  transform='Affine'):
    """Keypoint-supervised registration (reference _dock_widget.py:707-717): least-squares affine on the
    3xK keypoint pairs, then ICP on the full clouds.  Returns dict(transform, transform_sc, transform_icp,
    icp_residuals) as numpy arrays."""
    if transform != 'Affine':
        raise NotImplementedError("transform='Similar' is a SURVEY §8(f) 'next' row; only 'Affine' is built")
    m = D.to_device_points(moving)
    f = D.to_device_points(fixed)
    a_sc = D.fit_affine(D.to_device_points(moving_keypoints), D.to_device_points(fixed_keypoints))
    moved = D.apply_affine(m, a_sc)
    a_icp, resid, _ = D.icp_affine(moved, f, int(icp_iterations))
    a_final = D.compose(a_icp, a_sc)
    return _to_host(dict(transform=a_final, transform_sc=a_sc, transform_icp=a_icp, icp_residuals=resid))
