"""Drop-in for `platymatch/estimate_transform/shape_context.py` (reference :88-188).

get_unary / get_unary_distance / do_ransac keep the reference's names, argument meaning and return
conventions (numpy float64 in and out); the work runs on the GPU through the C ABI.
"""
import numpy as np

from .. import device as D

__all__ = ["get_unary", "get_unary_counts", "get_unary_distance", "unary_distance_matrix", "do_ransac"]


def get_unary_counts(centroid, mean_distance, detections, type, transposed=False):
    """Integer histograms behind get_unary: (counts [V,N,360] uint32, dropped [V,N], edge_ties, x0[3]).

    V = 2 for type='moving' (sc, sc2), 4 for type='fixed' (sc, sc2, sc3, sc4) — reference
    shape_context.py:170-185.  The reference row is counts / counts.sum() (:41).
    """
    torch = D._torch()
    pts = D.to_device_points(detections, transposed=transposed)
    c = torch.from_numpy(np.ascontiguousarray(np.asarray(centroid, dtype=np.float64).reshape(-1)[:3])).to(pts.device)
    md = torch.tensor([float(mean_distance)], dtype=torch.float64, device=pts.device)
    stats = D.cloud_stats(pts)                      # PCA first axis (:162-165)
    nvar = 4 if type == 'fixed' else 2
    counts, dropped, ties = D.shape_context_counts(pts, c, stats[3:6], md, nvar)
    return (counts.cpu().numpy().view(np.uint32), dropped.cpu().numpy().view(np.uint32), int(ties.item()),
            stats[3:6].cpu().numpy())


def get_unary(centroid, mean_distance, detections, type, transposed=False):
    """reference shape_context.py:144-188 — (sc, sc2, sc3, sc4), each (N,360) float64 rows summing to 1;
    sc3 / sc4 are empty arrays for type='moving' (:188)."""
    counts, _, _, _ = get_unary_counts(centroid, mean_distance, detections, type, transposed)
    out = []
    for v in range(4):
        if v < counts.shape[0]:
            c = counts[v].astype(np.float64)
            out.append(c / c.sum(axis=1, keepdims=True))          # sc / sc.sum()  (:41)
        else:
            out.append(np.array([]))
    return tuple(out)


def _hist_to_device(sc):
    """(N,360) normalised float64 histograms -> chi^2 operand on the GPU (float32 arithmetic)."""
    torch = D._torch()
    sc = np.ascontiguousarray(np.asarray(sc, dtype=np.float64).astype(np.float32))
    return D.chi2_operand(torch.from_numpy(sc).cuda())


def unary_distance_matrix(sc_a, sc_b):
    """The double loops of reference _dock_widget.py:556-602 as one call: U[i,j] =
    get_unary_distance(sc_a[i], sc_b[j]); float32 arithmetic on the GPU, returned as float64."""
    sc_a, sc_b = np.atleast_2d(sc_a), np.atleast_2d(sc_b)
    cost = D.chi2_cost(_hist_to_device(sc_a), _hist_to_device(sc_b))
    return cost[:, :sc_b.shape[0]].cpu().numpy().astype(np.float64)


def get_unary_distance(sc1, sc2):
    """reference shape_context.py:88-99 — chi^2 distance of two (360,) histograms (python float)."""
    return float(unary_distance_matrix(np.asarray(sc1).reshape(1, -1), np.asarray(sc2).reshape(1, -1))[0, 0])


def do_ransac(moving_all, fixed_all, min_samples=4, trials=500, error=5, transform='Affine', sample_indices=None,
              seed=None):
    """reference shape_context.py:103-139 — (A_best 4x4, inliers_best int).

    moving_all / fixed_all are 3xK (or 4xK) in correspondence order.  The reference draws samples
    from numpy's global unseeded RNG; here they come from a device Philox stream (`seed`, default:
    fresh entropy) or from an explicit (trials, min_samples) index array (`sample_indices`), which is
    what the parity tests use.
    """
    D.transform_code(transform)          # 'Affine' or 'Similar' (:126-129); anything else raises ValueError
    torch = D._torch()
    m = D.to_device_points(moving_all)
    f = D.to_device_points(fixed_all)
    idx = None
    if sample_indices is not None:
        idx = torch.from_numpy(np.ascontiguousarray(sample_indices, dtype=np.int32)).to(m.device)
        trials = idx.shape[0]
    if seed is None:
        seed = int(np.random.SeedSequence().entropy & (2 ** 63 - 1))
    a, inl, _, _ = D.ransac(m, f, int(trials), float(error), int(min_samples), idx, seed, transform=transform)
    return a.cpu().numpy().reshape(4, 4), int(inl.item())
