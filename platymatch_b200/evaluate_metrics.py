"""Matching accuracy and average registration error, the two numbers of the reference's `EvaluateMetrics`
widget (`_calculate_metrics`, reference _dock_widget.py:1030-1080).  SURVEY §8(f) row 3.

The three `cdist` + `linear_sum_assignment` pairs run on the GPU (pm_cdist + the exact assignment kernel of
the registration path); the dictionary bookkeeping on a handful of keypoints stays in Python as in the reference.
"""
import numpy as np

from . import device as D

__all__ = ["assign_by_distance", "calculate_metrics"]


def assign_by_distance(points_a, points_b):
    """linear_sum_assignment(cdist(a.T, b.T)) for 3 x Na / 3 x Nb clouds (:1032-1033, :1038-1039, :1050-1051):
    (row_ind ascending, col_ind) as scipy returns them."""
    torch = D._torch()
    a, b = D.to_device_points(points_a), D.to_device_points(points_b)
    na, nb = a.shape[0], b.shape[0]
    if na <= nb:
        cost = D.cdist(a, b)
        col4row, _, st = D.lap_solve(cost.unsqueeze(0), na, nb)
        rows, cols = np.arange(na, dtype=np.int64), col4row[0].cpu().numpy().astype(np.int64)
    else:                                   # scipy solves the transpose; pairs are reported by ascending row
        cost = D.cdist(b, a)
        col4row, _, st = D.lap_solve(cost.unsqueeze(0), nb, na)
        r = col4row[0].cpu().numpy().astype(np.int64)
        order = np.argsort(r)
        rows, cols = r[order], np.arange(nb, dtype=np.int64)[order]
    if int(st[0, 4].item()) != 0:
        raise ValueError("cost matrix is infeasible")
    return rows, cols


def calculate_metrics(moving_keypoints, moving_keypoint_ids, moving_detections, moving_ids,
                      fixed_keypoints, fixed_keypoint_ids, fixed_detections, fixed_ids,
                      transform_matrix_1, transform_matrix_2=None):
    """reference _dock_widget.py:1030-1080 -> (matching_accuracy, average_registration_error) as floats.

    keypoints / detections are 3 x K and 3 x N (zyx) with their id vectors; the moving cloud is mapped through
    transform_matrix_1 and then transform_matrix_2 (identity if omitted; :1044-1046, combined at :1028).
    """
    from .estimate_transform.apply_transform import apply_affine_transform
    moving_keypoint_ids, fixed_keypoint_ids = np.asarray(moving_keypoint_ids), np.asarray(fixed_keypoint_ids)
    moving_ids, fixed_ids = np.asarray(moving_ids), np.asarray(fixed_ids)
    t1 = np.asarray(transform_matrix_1, dtype=np.float64)
    t2 = np.eye(4) if transform_matrix_2 is None else np.asarray(transform_matrix_2, dtype=np.float64)
    # first associate keypoints with detections (:1032-1042)
    r, c = assign_by_distance(moving_keypoints, moving_detections)
    moving_dictionary = {moving_keypoint_ids[i]: moving_ids[c[k]] for k, i in enumerate(r)}
    r, c = assign_by_distance(fixed_keypoints, fixed_detections)
    fixed_dictionary = {fixed_keypoint_ids[i]: fixed_ids[c[k]] for k, i in enumerate(r)}
    # transformed moving detections against fixed detections (:1044-1051)
    moved = apply_affine_transform(apply_affine_transform(moving_detections, t1), t2)
    row_indices, col_indices = assign_by_distance(moved, fixed_detections)
    row_ids, col_ids = moving_ids[row_indices], fixed_ids[col_indices]
    hits = 0
    for key in moving_dictionary.keys():                                            # :1058-1062
        if key in fixed_dictionary.keys():
            if np.any(col_ids[np.where(row_ids == moving_dictionary[key])] == fixed_dictionary[key]):
                hits += 1
    accuracy = hits / len(fixed_dictionary.keys())                                  # :1067
    combined = np.matmul(t2, t1)                                                    # :1028
    moved_kp = apply_affine_transform(moving_keypoints, combined)
    distance = 0.0
    fk = np.asarray(fixed_keypoints, dtype=np.float64)[:3].transpose()
    for i in range(moved_kp.shape[1]):                                              # :1072-1076
        sel = fk[np.where(fixed_keypoint_ids == moving_keypoint_ids[i])]
        distance += np.linalg.norm(sel - moved_kp.transpose()[i, :])
    return accuracy, distance / len(moving_dictionary.keys())                        # :1079
