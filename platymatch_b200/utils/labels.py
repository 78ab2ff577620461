"""Label image -> detections, the first stage of `EstimateTransform._click_run` when the inputs are
instance-segmentation volumes instead of CSV detections (reference _dock_widget.py:497-521), and the
data-dependent RANSAC threshold derived from the nucleus sizes (:613-618).  SURVEY §8(f) row 2.
"""
import numpy as np

from .. import device as D

__all__ = ["detections_from_labels", "ransac_error_from_sizes"]


def detections_from_labels(label_image, anisotropy=1.0):
    """reference _dock_widget.py:497-509 (moving) / :512-521 (fixed).

    label_image: (nz, ny, nx) integer array, 0 = background.  Returns
      detections  (3, N) float64, rows z, y, x = np.mean of the voxel indices of each id (:506-507,520-521)
      sizes       (N,)  float64 = anisotropy * voxel count (:508,517)
      ids         (N,)  the non-zero ids in np.unique order (:497-500)
    One streaming pass on the GPU instead of one np.where pass per id.
    """
    torch = D._torch()
    a = np.asarray(label_image)
    if a.ndim != 3:
        raise ValueError("expected a 3-D label volume, got %d-D" % a.ndim)
    if not np.issubdtype(a.dtype, np.integer):
        raise ValueError("label volume must have an integer dtype")
    if a.dtype == np.uint16:
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).cuda().view(torch.uint16)
    else:
        if a.size and (a.max() > np.iinfo(np.int32).max):
            raise ValueError("label ids above 2^31 - 1 are not supported")
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda()
    ids, cen, sizes = D.label_centroids(t, anisotropy)
    return cen.cpu().numpy().T.copy(), sizes.cpu().numpy(), ids.cpu().numpy().astype(a.dtype, copy=False)


def ransac_error_from_sizes(moving_nucleus_size, fixed_nucleus_size):
    """reference _dock_widget.py:613-618 — 16 without sizes (CSV mode), else the approximate mean nucleus radius."""
    if moving_nucleus_size is None or fixed_nucleus_size is None or len(moving_nucleus_size) == 0 \
            or len(fixed_nucleus_size) == 0:
        return 16
    return 0.5 * (np.average(moving_nucleus_size) ** (1 / 3) + np.average(fixed_nucleus_size) ** (1 / 3))
