"""Drop-in for `platymatch/estimate_transform/find_transform.py` (reference :4-17)."""
import numpy as np

from .. import device as D

__all__ = ["get_affine_transform", "get_similar_transform"]


def get_affine_transform(moving, fixed, with_ones=False):
    """reference find_transform.py:4-17 — fixed_h @ pinv(moving_h) for 3xK clouds (K >= 4), 4x4 out.

    Evaluated as the float64 normal equations F M^T (M M^T)^-1 on the GPU, which equals the
    pseudo-inverse solution for full-rank (non-coplanar) point sets; rank-deficient inputs return NaN
    rows instead of numpy's minimum-norm answer.
    """
    moving = np.asarray(moving, dtype=np.float64)
    fixed = np.asarray(fixed, dtype=np.float64)
    if with_ones:
        moving, fixed = moving[:3], fixed[:3]
    a = D.fit_affine(D.to_device_points(moving), D.to_device_points(fixed))
    return a.cpu().numpy().reshape(4, 4)


def get_similar_transform(moving, fixed):
    """reference find_transform.py:21-99 (Horn's closed form).  SURVEY.md §8(f) row 1 — "next": the
    affine mode is the north-star path; the similarity mode is not built yet."""
    raise NotImplementedError("transform='Similar' is a SURVEY §8(f) 'next' row; only 'Affine' is built")
