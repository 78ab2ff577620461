"""Drop-in for `platymatch/estimate_transform/apply_transform.py` (reference :3-17)."""
import numpy as np

from .. import device as D

__all__ = ["apply_affine_transform"]


def apply_affine_transform(moving, affine_transform_matrix):
    """reference apply_transform.py:3-17 — moving 3xN (or 4xN), 4x4 matrix -> 3xN."""
    torch = D._torch()
    pts = D.to_device_points(moving, transposed=False)
    a = torch.from_numpy(np.ascontiguousarray(affine_transform_matrix, dtype=np.float64).reshape(16)).to(pts.device)
    return np.ascontiguousarray(D.apply_affine(pts, a).cpu().numpy().T)
