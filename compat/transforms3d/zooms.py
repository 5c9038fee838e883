import numpy as np


def zfdir2mat(factor, direction=None):
    """3x3 zoom by `factor`: isotropic if direction is None, else along `direction` only
    (M = I + (factor - 1) * d d^T for the unit vector d)."""
    if direction is None:
        return np.eye(3) * factor
    d = np.asarray(direction, dtype=np.float64)
    d = d / np.linalg.norm(d)
    return np.eye(3) + (factor - 1.0) * np.outer(d, d)
