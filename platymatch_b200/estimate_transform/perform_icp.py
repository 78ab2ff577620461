"""Drop-in for `platymatch/estimate_transform/perform_icp.py` (reference :7-26)."""
import numpy as np

from .. import device as D

__all__ = ["perform_icp"]


def perform_icp(moving, fixed, icp_iterations=50, transform='Affine', verbose=True, return_residuals=False):
    """reference perform_icp.py:7-26 — 3xN clouds, exactly `icp_iterations` iterations, 4x4 out.

    The reference prints the mean residual of every iteration (:24); `verbose` keeps that behaviour.
    """
    m = D.to_device_points(moving)
    f = D.to_device_points(fixed)
    a_icp, resid, _ = D.icp(m, f, int(icp_iterations), transform=transform)       # 'Affine' | 'Similar' (:17-20)
    resid = resid.cpu().numpy()
    if verbose:
        for i, r in enumerate(resid):
            print("Residual at iteration {} is {}".format(str(i), r))
    a = a_icp.cpu().numpy().reshape(4, 4)
    return (a, resid) if return_residuals else a
