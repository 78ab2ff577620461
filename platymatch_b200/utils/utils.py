"""Drop-in for the numeric half of `platymatch/utils/utils.py` (reference :48-88), numpy in / numpy out.

Same names, argument meaning and return shapes as the reference; the arithmetic runs on the GPU
through the C ABI (no CPU fallback).
"""
import numpy as np

from .. import device as D

__all__ = ["get_centroid", "get_mean_distance", "get_error"]


def get_centroid(detections, transposed=True):
    """reference utils.py:48-56 — (1,3) if transposed (N x 3/4 input) else (3,1) (3/4 x N input)."""
    pts = D.to_device_points(detections, transposed=transposed)
    c = D.cloud_stats(pts)[0:3].cpu().numpy()
    return c.reshape(1, 3) if transposed else c.reshape(3, 1)


def get_mean_distance(detections, transposed=True):
    """reference utils.py:58-75 — mean Euclidean distance over all unordered pairs (python float)."""
    pts = D.to_device_points(detections, transposed=transposed)
    return float(D.mean_distance(pts).cpu().numpy()[0])


def get_error(moving_landmarks, fixed_landmarks):
    """reference utils.py:77-88 — mean residual norm of 3 x N landmark sets (host arithmetic: O(N))."""
    if moving_landmarks is not None or fixed_landmarks is not None:
        residual = np.asarray(moving_landmarks) - np.asarray(fixed_landmarks)
        return np.mean(np.linalg.norm(residual, axis=0))
    return None
