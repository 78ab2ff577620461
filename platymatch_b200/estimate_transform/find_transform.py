"""Drop-in for `platymatch/estimate_transform/find_transform.py` (reference :4-17, :21-99)."""
import numpy as np

from .. import device as D

__all__ = ["get_affine_transform", "get_similar_transform"]


def get_affine_transform(moving, fixed, with_ones=False):
    """reference find_transform.py:4-17 — fixed_h @ pinv(moving_h) for 3xK clouds, 4x4 out.

    Evaluated as the float64 normal equations F M^T (M M^T)^+ on the GPU: Gauss-Jordan for full-rank
    (non-coplanar) point sets, and numpy's minimum-norm pseudo-inverse answer for rank-deficient ones
    (keypoints picked in one z slice, collinear points, K < 4) through a 4x4 eigen-decomposition.
    """
    moving = np.asarray(moving, dtype=np.float64)
    fixed = np.asarray(fixed, dtype=np.float64)
    if with_ones:
        moving, fixed = moving[:3], fixed[:3]
    a = D.fit_transform(D.to_device_points(moving), D.to_device_points(fixed), 'Affine')
    return a.cpu().numpy().reshape(4, 4)


def get_similar_transform(moving, fixed):
    """reference find_transform.py:21-99 — scale x rotation + translation by Horn's closed form, 4x4 out.

    Built as PUBLISHED (the unit quaternion is the eigenvector of the largest eigenvalue of Horn's 4x4 matrix).
    The reference as shipped takes `q = D[0]` (:66), the first ROW of numpy's eigenvector matrix — the first
    components of four different eigenvectors with LAPACK's arbitrary signs — which is not a rotation and cannot
    be reproduced by another eigen-solver; see DESIGN.md "Similar" (goldens of both variants are kept in
    tests/golden/similar.npz).
    """
    moving = np.asarray(moving, dtype=np.float64)[:3]
    fixed = np.asarray(fixed, dtype=np.float64)[:3]
    a = D.fit_transform(D.to_device_points(moving), D.to_device_points(fixed), 'Similar')
    return a.cpu().numpy().reshape(4, 4)
