"""The reference's file formats: detections CSV and 4 x 4 transform text files.

reference platymatch/utils/utils.py:19-34 (`_browse_detections`), :37-43 (`_browse_transform`) and
platymatch/_dock_widget.py:426-432 (`_save_transform`), without the Qt file dialogs.
"""
import numpy as np

__all__ = ["load_detections", "load_transform", "save_transform"]


def load_detections(path, header=False, izyx=False):
    """Space-delimited rows `id x y z [...]` (utils.py:22-27); the first row is skipped when `header` (:23).
    Columns 1:4 are flipped to z y x unless the file is already `id z y x r` (`izyx`, :30-33).
    Returns (detections 3 x N float64, ids N) like the reference (:34)."""
    rows = np.loadtxt(path, delimiter=" ", skiprows=1 if header else 0, ndmin=2)
    ids = rows[:, 0]
    det = rows[:, 1:4].astype(np.float64)
    if not izyx:
        det = np.flip(det, 1)
    return np.ascontiguousarray(det.transpose()), ids.transpose()


def load_transform(path):
    """4 x 4 space-delimited matrix (utils.py:37-43)."""
    a = np.loadtxt(path, delimiter=" ", ndmin=2).astype(np.float64)
    assert a.shape == (4, 4), 'Loaded transform does not have shape 4 x 4'
    return a


def save_transform(path, transform_matrix_icp, transform_matrix_sc=None):
    """np.savetxt(icp @ sc, delimiter=' ', fmt='%1.3f') (_dock_widget.py:428-432); pass one matrix to save it as is."""
    a = np.asarray(transform_matrix_icp, dtype=np.float64)
    if transform_matrix_sc is not None:
        a = np.matmul(a, np.asarray(transform_matrix_sc, dtype=np.float64))
    np.savetxt(path, a, delimiter=' ', fmt='%1.3f')
    return a
